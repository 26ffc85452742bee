"""The C-ABI library loads and exports exactly what include/maxk_b200.h declares (no GPU)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from spgemm_gnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "maxk_b200.h")).read()
    return sorted(set(re.findall(r"MK_API\s+[\w\s\*]+?\b(mk_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built_lib):
    names = _declared()
    assert len(names) >= 10
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table covers the header
    L = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib]).decode()
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == names                          # and nothing else leaks out


def test_header_is_plain_c_and_links(built_lib, tmp_path):
    """A C99 host (tests/c/abi_smoke.c) compiles against include/maxk_b200.h, links the library and
    gets the documented return codes -- the boundary needs neither Python nor C++."""
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe,
                           "-L", libdir, "-lmaxk_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode()
    assert "c abi ok: version 202" in out


def test_library_is_sm100a_and_has_no_torch_dependency(built_lib):
    out = subprocess.check_output(["cuobjdump", "-lelf", built_lib]).decode()
    assert "sm_100a" in out and "sm_80" not in out
    dyn = subprocess.check_output(["readelf", "-d", built_lib]).decode()
    needed = " ".join(l for l in dyn.splitlines() if "NEEDED" in l)
    assert "torch" not in needed and "c10" not in needed and "libcudart" not in needed


def _sass(built_lib, mangled_prefix):
    """SASS text of every kernel of the library whose mangled name starts with the prefix."""
    names = subprocess.check_output(["cuobjdump", "-elf", built_lib]).decode()
    import re
    funs = sorted(set(re.findall(r"\.text\.(" + re.escape(mangled_prefix) + r"\w*)", names)))
    assert funs, mangled_prefix
    return "".join(subprocess.check_output(["cuobjdump", "-sass", "-fun", f, built_lib]).decode() for f in funs[:1])


def test_sass_carries_the_instructions_the_design_names(built_lib):
    """DESIGN.md section 5 by mnemonic: the lane-contiguous top-k loads rows with 256-bit loads, bounds the search
    with a warp minimum and stages entries with 64-bit shared stores; the backward reduces with four-wide float
    reductions in L2; the forward never uses a float atomic; the multicast exchange uses the switch."""
    topk = _sass(built_lib, "_ZN2mk21topk_cbsr_lane_kernelILi2EhLi8ELb1E")
    assert "LDG.E.NA.ENL2.256" in topk or ".256" in topk
    assert "CREDUX.MIN" in topk and "STS.64" in topk and "REDUX.SUM" in topk
    bwd = _sass(built_lib, "_ZN2mk22sspmm_bwd_kernel_occ32ILi32Eh")
    assert "REDG.E.ADD.F32x4" in bwd or "RED.E.ADD.F32x4" in bwd
    fwd = _sass(built_lib, "_ZN2mk24spgemm_fwd_banked_kernelILi32ELi4ELb0ELb0E")
    assert "ATOM" not in fwd and "RED." not in fwd and "LDS" in fwd and "STS" in fwd
    mc = _sass(built_lib, "_ZN2mk29peer_reduce_scatter_mc_kernel")
    assert "LDGMC" in mc and "ADD.F32x4" in mc


def test_version_and_error_strings(built_lib):
    L = _lib.lib()
    assert L.mk_version() == 202
    assert L.mk_error_string(0) == b"ok"
    assert L.mk_error_string(-1) == b"invalid argument"
    assert b"CUDA" in L.mk_error_string(-3)


def test_argument_validation_without_a_device(built_lib):
    """Bad arguments are rejected before any CUDA call, so this runs on the CPU box."""
    L = _lib.lib()
    assert L.mk_topk_cbsr(None, 4, 16, 0, None, None, 1, None) == _lib.MK_EINVAL      # k < 1
    assert L.mk_topk_cbsr(None, 4, 16, 17, None, None, 1, None) == _lib.MK_EINVAL     # k > d
    assert L.mk_topk_cbsr(None, 4, 300, 8, None, None, 1, None) == _lib.MK_EINVAL     # u8 with d>256
    assert L.mk_topk_cbsr(None, 4, 16, 8, None, None, 3, None) == _lib.MK_EINVAL      # index_bytes
    assert L.mk_topk_cbsr(None, 0, 16, 8, None, None, 1, None) == _lib.MK_OK          # empty
    assert L.mk_spgemm_fwd(None, 0, 0, None, None, None, None, 1, None, None, 0, 8, 16, None) == _lib.MK_OK
    assert L.mk_spgemm_fwd(None, 5, 0, None, None, None, None, 1, None, None, 5, 8, 4, None) == _lib.MK_EINVAL
    assert L.mk_partition(None, -1, 64, None, None, None, None) == _lib.MK_EINVAL
    assert L.mk_partition(None, 10, 0, None, None, None, None) == _lib.MK_EINVAL


def test_peer_argument_validation_without_a_device(built_lib):
    """The peer-exchange entry points reject bad windows / offsets before any CUDA call."""
    L = _lib.lib()
    VP = ctypes.c_void_p
    wins = (VP * 16)()
    wins[0], wins[1] = 0x1000, 0x2000
    nb, off = (ctypes.c_int64 * 1)(64), (ctypes.c_int64 * 1)(1024)
    E = _lib.MK_EINVAL
    assert L.mk_peer_push(wins, 0, 0, 1, off, nb, None) == E          # world < 1
    assert L.mk_peer_push(wins, 17, 0, 1, off, nb, None) == E         # world > 16
    assert L.mk_peer_push(wins, 2, 2, 1, off, nb, None) == E          # rank >= world
    assert L.mk_peer_push(wins, 3, 0, 1, off, nb, None) == E          # window 2 missing
    assert L.mk_peer_push(wins, 2, 0, 9, off, nb, None) == E          # > 8 segments
    off[0] = 512
    assert L.mk_peer_push(wins, 2, 0, 1, off, nb, None) == E          # inside the header
    off[0] = 1024
    assert L.mk_peer_push(wins, 1, 0, 1, off, nb, None) == _lib.MK_OK  # one rank: nothing to send
    assert L.mk_peer_begin_push(None, 2, 0, 0, 0, None) == E           # no window
    assert L.mk_peer_begin_push(0x1000, 2, 0, 2, 0, None) == E         # only two table buffers
    assert L.mk_peer_publish(None, 0, 0, None) == E
    assert L.mk_peer_publish(0x1000, 0, 3, None) == E
    assert L.mk_peer_wait_all(None, 2, 0, None) == E
    assert L.mk_peer_release(wins, 2, 5, None) == E                    # rank >= world
    assert L.mk_peer_reduce_scatter(wins, 2, 0, 1000, 64, None, 0, 0, None) == E     # offset in header
    assert L.mk_peer_reduce_scatter(wins, 2, 0, 1024, 64, None, 0, 0, None) == E     # no output
    assert L.mk_peer_reduce_scatter(wins, 2, 0, 1024, 24, None, 0, 0, None) == E     # not 16-byte units
    # waiting / pushing forward: a bad exchange description is rejected before any CUDA call
    x = _lib.FwdExchange()
    x.window, x.world, x.rank, x.rows_per_rank = 0x1000, 2, 2, 100
    fwd = lambda: L.mk_spgemm_fwd_banked_ex(None, 5, 0, None, None, None, None, None, None, None, 5, 32, 256,
                                            None, ctypes.byref(x), None)
    assert fwd() == E                                                  # rank >= world
    x.rank, x.rows_per_rank = 0, 0
    assert fwd() == E                                                  # no rows per rank
    x.rows_per_rank, x.pushers, x.n_seg = 100, 4, 1
    x.h_windows, x.h_offsets, x.h_bytes = ctypes.cast(wins, ctypes.POINTER(VP)), off, nb
    nb[0] = 24
    assert fwd() == E                                                  # segment not in 16-byte units
    nb[0] = 64
    x.window = 0x5000
    assert fwd() == E                                                  # h_windows[rank] is not the window
    assert L.mk_peer_push_sm(wins, 2, 0, 1, off, nb, 0, None) == E     # no pushers
    assert L.mk_peer_push_sm(wins, 2, 0, 4, off, nb, 8, None) == E     # > 3 segments
    p = VP(0)
    assert L.mk_peer_alloc(16, ctypes.byref(p)) == E                                # smaller than the header
    assert L.mk_peer_free(None) == _lib.MK_OK and L.mk_peer_close(None) == _lib.MK_OK


def test_peer_layout_and_gating():
    from spgemm_gnn_b200 import peer
    offs, total = peer.layout([1000, 4096, 1])
    assert offs == [1024, 2048, 6144] and total == 6400
    assert all(o % 256 == 0 for o in offs) and total % 256 == 0
    assert not peer.available()          # no NCCL process group here: dist.py stays on its collectives


def test_entry_points_refuse_cpu_tensors(built_lib):
    """TORCH_CHECK messages of the reference binding."""
    import maxk_kernels as mk
    x = torch.randn(4, 16)
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        mk.maxk_forward(x, 4)
    with pytest.raises(RuntimeError, match="grad_output must be a CUDA tensor"):
        mk.maxk_backward(x, torch.zeros(4, 16, dtype=torch.uint8))
    ptr = torch.zeros(5, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="ptr must be a CUDA tensor"):
        mk.spgemm_forward(ptr, ptr, x, x, x, 4, 0, 4, 16)
    with pytest.raises(RuntimeError, match="ptr must be a CUDA tensor"):
        mk.spgemm_backward(ptr, ptr, x, x, x, 4, 0, 4, 16)
    assert mk.__all__[:4] == ["maxk_forward", "maxk_backward", "spgemm_forward", "spgemm_backward"]


def test_product_never_imports_the_oracle():
    """The shipped path must not route through oracle/ (checked textually)."""
    bad = []
    for base in ("spgemm_gnn_b200", "."):
        d = os.path.join(ROOT, base)
        for f in os.listdir(d):
            if f.endswith(".py") and f not in ("bench.py", "__graft_entry__.py"):
                src = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MaxKLibraryError, match="no CPU fallback"):
        _lib.lib()
