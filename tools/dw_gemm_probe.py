"""Weight-gradient GEMM of the layers' Linear (dW = dZ^T X, reduction over 232,965 nodes, 256 x 256 / 256 x 608 output):
what torch's autograd calls against equivalent formulations (TF32, as the epoch measurement)."""
import torch

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
n = 232965
gen = torch.Generator(device="cuda").manual_seed(1)


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for kin in (256, 608):
    x = torch.randn(n, kin, device="cuda", generator=gen)
    dz = torch.randn(n, 256, device="cuda", generator=gen)
    w = torch.randn(256, kin, device="cuda", generator=gen)
    ref = dz.t().mm(x)
    forms = {
        "dz.t().mm(x)  [autograd]": lambda: dz.t().mm(x),
        "x.t().mm(dz).t()": lambda: x.t().mm(dz).t(),
        "einsum nk,nj->kj": lambda: torch.einsum("nk,nj->kj", dz, x),
        "8 row chunks, summed": lambda: sum(dz[i::8].t().mm(x[i::8]) for i in range(8)),
        "contiguous chunks via baddbmm": lambda: torch.bmm(dz[: n // 5 * 5].view(5, n // 5, 256).transpose(1, 2),
                                                            x[: n // 5 * 5].view(5, n // 5, kin)).sum(0),
    }
    for name, fn in forms.items():
        out = fn()
        err = float((out - ref).abs().max() / ref.abs().max()) if out.shape == ref.shape else float("nan")
        print(f"K_in={kin}: {name}: {t(fn):.4f} ms (max rel diff to autograd form {err:.1e})", flush=True)
    print(f"K_in={kin}: forward x @ w.t(): {t(lambda: x.mm(w.t())):.4f} ms; dX = dz @ w: {t(lambda: dz.mm(w)):.4f} ms", flush=True)


def dw_splitk(gy, x, chunks):
    n = gy.shape[0]
    rows = n // chunks
    main = rows * chunks
    out = torch.bmm(gy[:main].view(chunks, rows, gy.shape[1]).transpose(1, 2), x[:main].view(chunks, rows, x.shape[1])).sum(0)
    if main < n:
        out += gy[main:].t().mm(x[main:])
    return out


print("--- chunk count of the batched form (remainder rows in a mm of their own)")
for nn_, kin in ((232965, 256), (232965, 608), (2449029, 256), (89250, 256), (29121, 256), (232965, 48)):
    x = torch.randn(nn_, kin, device="cuda", generator=gen)
    dz = torch.randn(nn_, 256, device="cuda", generator=gen)
    base = t(lambda: dz.t().mm(x))
    row = [f"autograd form {base:.4f}"]
    for ch in (4, 8, 16, 32, 64, 128):
        row.append(f"{ch}: {t(lambda: dw_splitk(dz, x, ch)):.4f}")
    print(f"n={nn_} K_in={kin}: " + "  ".join(row), flush=True)
    del x, dz
