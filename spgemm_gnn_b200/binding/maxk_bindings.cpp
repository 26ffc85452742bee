// `maxk_kernels` as a compiled pybind11 / ATen extension -- the form the reference ships
// (kernels/maxk_bindings.cpp is absent from the snapshot; setup.py:25-31 names it, the four entry
// points and their TORCH_CHECK strings are recovered from the binary):
//
//     maxk_forward(input, k) -> Tensor                      [N, k] kept values
//     maxk_backward(grad_output, indices) -> Tensor         [N, D] dense gradient
//     spgemm_forward(ptr, idx, val, sp_data, sp_index, num_nodes, num_edges, dim_sparse, dim_origin)
//         -> (Tensor out [num_nodes, dim_origin], Tensor sp_index)
//     spgemm_backward(ptr, idx, val, grad_output, sp_index, num_nodes, num_edges, dim_sparse, dim_origin)
//         -> Tensor [sp_index.size(0), dim_sparse]
//
// plus maxk_forward_cbsr(input, k) -> (sp_data, sp_index).  Every function is a few checks and
// one or two calls into the C ABI of libmaxk_b200.so (include/maxk_b200.h) on torch's current CUDA
// stream, with the GIL released; there is no kernel code and no CPU path in this file.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <mutex>
#include <tuple>
#include <unordered_map>

#include "maxk_b200.h"

namespace {

void chk(int rc, const char* what) {
    if (rc == MK_OK) return;
    TORCH_CHECK(false, what, " failed: ", mk_error_string(rc), rc == MK_ECUDA ? ": " : "",
                rc == MK_ECUDA ? mk_last_cuda_error() : "");
}

void* stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void cuda_contig(const torch::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

int index_bytes(const torch::Tensor& sp_index, int64_t dim_origin) {
    if (sp_index.scalar_type() == torch::kUInt8) {
        TORCH_CHECK(dim_origin <= 256, "sp_index must be uint16 when dim_origin > 256");
        return 1;
    }
    TORCH_CHECK(sp_index.scalar_type() == torch::kUInt16 || sp_index.scalar_type() == torch::kInt16,
                "sp_index must be uint8 or uint16");
    return 2;
}

// ---- work records per graph: built once on the GPU, kept while the row pointer lives ----------
struct Records {
    torch::Tensor parts, exec;  // row order (fold), longest first (what the CTAs take)
    int64_t num_parts = 0, num_slots = 0;
    uint32_t version = 0;
};
std::mutex g_mu;
std::unordered_map<uint64_t, Records> g_records;

Records& records_for(const torch::Tensor& ptr, int64_t num_nodes) {
    const uint64_t key = reinterpret_cast<uint64_t>(ptr.data_ptr()) ^ (static_cast<uint64_t>(num_nodes) << 48) ^
                         (static_cast<uint64_t>(ptr.get_device()) << 40);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_records.find(key);
    if (it != g_records.end() && it->second.version == ptr._version()) return it->second;
    if (g_records.size() > 64) g_records.clear();
    Records r;
    const int max_nz = 1024;
    int64_t np = 0, ns = 0;
    chk(mk_partition(ptr.data_ptr<int32_t>(), num_nodes, max_nz, nullptr, &np, &ns, stream()), "mk_partition");
    r.num_parts = np;
    r.num_slots = ns;
    r.parts = torch::empty({std::max<int64_t>(np, 1), 4}, ptr.options().dtype(torch::kInt32));
    chk(mk_partition(ptr.data_ptr<int32_t>(), num_nodes, max_nz, reinterpret_cast<mk_part*>(r.parts.data_ptr<int32_t>()),
                     nullptr, nullptr, stream()),
        "mk_partition");
    r.exec = r.parts;
    if (np > 1) {
        auto lens = r.parts.slice(0, 0, np).select(1, 2);
        if (lens.sum().item<int64_t>() >= 96 * np)  // long records: longest first
            r.exec = r.parts.slice(0, 0, np).index_select(0, lens.argsort(/*stable=*/true, 0, /*descending=*/true)).contiguous();
    }
    r.version = ptr._version();
    return g_records[key] = std::move(r);
}

void check_graph(const torch::Tensor& ptr, const torch::Tensor& idx, const torch::Tensor& val) {
    cuda_contig(ptr, "ptr");
    cuda_contig(idx, "idx");
    cuda_contig(val, "val");
    TORCH_CHECK(ptr.scalar_type() == torch::kInt32, "ptr must be int32");
    TORCH_CHECK(idx.scalar_type() == torch::kInt32, "idx must be int32");
    TORCH_CHECK(val.scalar_type() == torch::kFloat32, "val must be float32");
}

}  // namespace

std::tuple<torch::Tensor, torch::Tensor> maxk_forward_cbsr(torch::Tensor input, int64_t k) {
    cuda_contig(input, "input");
    TORCH_CHECK(input.dim() == 2, "Input must be 2D tensor");
    TORCH_CHECK(input.scalar_type() == torch::kFloat32, "input must be float32");
    const int64_t n = input.size(0), d = input.size(1);
    TORCH_CHECK(k >= 1 && k <= d, "k must be between 1 and input dimension");
    TORCH_CHECK(d <= 65536, "input dimension above 65536 is not supported");
    const c10::cuda::CUDAGuard guard(input.device());
    auto sp_data = torch::empty({n, k}, input.options());
    auto sp_index = torch::empty({n, k}, input.options().dtype(d <= 256 ? torch::kUInt8 : torch::kUInt16));
    chk(mk_topk_cbsr(input.data_ptr<float>(), n, static_cast<int>(d), static_cast<int>(k), sp_data.data_ptr<float>(),
                     sp_index.data_ptr(), d <= 256 ? 1 : 2, stream()),
        "mk_topk_cbsr");
    return {sp_data, sp_index};
}

torch::Tensor maxk_forward(torch::Tensor input, int64_t k) { return std::get<0>(maxk_forward_cbsr(input, k)); }

torch::Tensor maxk_backward(torch::Tensor grad_output, torch::Tensor indices) {
    cuda_contig(grad_output, "grad_output");
    cuda_contig(indices, "indices");
    TORCH_CHECK(grad_output.dim() == 2, "grad_output must be 2D tensor");
    TORCH_CHECK(grad_output.scalar_type() == torch::kFloat32, "grad_output must be float32");
    TORCH_CHECK(indices.sizes() == grad_output.sizes(), "indices must have the shape of grad_output");
    const int64_t n = grad_output.size(0), k = grad_output.size(1);
    // like the reference: the dense width is what the indices span (one device sync)
    int64_t d = indices.numel() ? indices.max().item<int64_t>() + 1 : 1;
    d = std::max(d, k);
    const c10::cuda::CUDAGuard guard(grad_output.device());
    if (indices.scalar_type() != torch::kUInt8 && indices.scalar_type() != torch::kUInt16 &&
        indices.scalar_type() != torch::kInt16)
        indices = d <= 256 ? indices.to(torch::kUInt8) : indices.to(torch::kInt16);
    auto dense = torch::empty({n, d}, grad_output.options());
    chk(mk_cbsr_scatter(grad_output.data_ptr<float>(), indices.data_ptr(), index_bytes(indices, d),
                        dense.data_ptr<float>(), n, static_cast<int>(k), static_cast<int>(d), stream()),
        "mk_cbsr_scatter");
    return dense;
}

std::tuple<torch::Tensor, torch::Tensor> spgemm_forward(torch::Tensor ptr, torch::Tensor idx, torch::Tensor val,
                                                        torch::Tensor sp_data, torch::Tensor sp_index,
                                                        int64_t num_nodes, int64_t num_edges, int64_t dim_sparse,
                                                        int64_t dim_origin) {
    check_graph(ptr, idx, val);
    cuda_contig(sp_data, "sp_data");
    cuda_contig(sp_index, "sp_index");
    TORCH_CHECK(sp_data.scalar_type() == torch::kFloat32, "sp_data must be float32");
    TORCH_CHECK(sp_data.dim() == 2 && sp_index.sizes() == sp_data.sizes(), "sp_index must have the shape of sp_data");
    TORCH_CHECK(sp_data.size(1) == dim_sparse, "dim_sparse must equal sp_data.size(1)");
    TORCH_CHECK(dim_sparse >= 1 && dim_sparse <= dim_origin, "k must be between 1 and input dimension");
    TORCH_CHECK(idx.numel() >= num_edges && val.numel() >= num_edges, "idx/val must hold num_edges entries");
    TORCH_CHECK(ptr.numel() >= num_nodes + 1, "ptr must have num_nodes + 1 entries");
    const int ib = index_bytes(sp_index, dim_origin);
    const c10::cuda::CUDAGuard guard(sp_data.device());
    Records& rec = records_for(ptr, num_nodes);
    const int k = static_cast<int>(dim_sparse), d = static_cast<int>(dim_origin);
    const int64_t n_src = sp_data.size(0);
    auto out = torch::empty({num_nodes, dim_origin}, sp_data.options());
    torch::Tensor partial;
    if (rec.num_slots > 0) partial = torch::empty({rec.num_slots, dim_origin}, sp_data.options());
    float* pp = rec.num_slots > 0 ? partial.data_ptr<float>() : nullptr;
    const mk_part* parts = reinterpret_cast<const mk_part*>(rec.parts.data_ptr<int32_t>());
    const mk_part* exec = rec.exec.is_same(rec.parts) ? nullptr : reinterpret_cast<const mk_part*>(rec.exec.data_ptr<int32_t>());
    const bool long_records = num_edges >= 96 * std::max<int64_t>(rec.num_parts, 1);
    py::gil_scoped_release nogil;
    if (long_records && k >= 32 && mk_banked_supported(k, d)) {
        auto bk_data = torch::empty_like(sp_data);
        auto bk_slot = torch::empty({n_src, dim_sparse}, sp_data.options().dtype(torch::kInt16));
        chk(mk_cbsr_bank(sp_data.data_ptr<float>(), sp_index.data_ptr(), ib, bk_data.data_ptr<float>(), nullptr,
                         reinterpret_cast<uint16_t*>(bk_slot.data_ptr<int16_t>()), n_src, k, d, stream()),
            "mk_cbsr_bank");
        chk(mk_spgemm_fwd_banked_ex(parts, rec.num_parts, rec.num_slots, exec, idx.data_ptr<int32_t>(),
                                    val.data_ptr<float>(), bk_data.data_ptr<float>(),
                                    reinterpret_cast<const uint16_t*>(bk_slot.data_ptr<int16_t>()), out.data_ptr<float>(),
                                    pp, num_nodes, k, d, nullptr, nullptr, stream()),
            "mk_spgemm_fwd_banked_ex");
    } else if (long_records && mk_packed_supported(k, d)) {
        auto pack = torch::empty({n_src, dim_sparse, 2}, sp_data.options().dtype(torch::kInt32));
        chk(mk_cbsr_bank_packed(sp_data.data_ptr<float>(), sp_index.data_ptr(), ib, pack.data_ptr(), n_src, k, d, stream()),
            "mk_cbsr_bank_packed");
        chk(mk_spgemm_fwd_packed_ex(parts, rec.num_parts, rec.num_slots, exec, idx.data_ptr<int32_t>(),
                                    val.data_ptr<float>(), pack.data_ptr(), out.data_ptr<float>(), pp, num_nodes, k, d,
                                    nullptr, nullptr, stream()),
            "mk_spgemm_fwd_packed_ex");
    } else {
        chk(mk_spgemm_fwd(parts, rec.num_parts, rec.num_slots, idx.data_ptr<int32_t>(), val.data_ptr<float>(),
                          sp_data.data_ptr<float>(), sp_index.data_ptr(), ib, out.data_ptr<float>(), pp, num_nodes, k, d,
                          stream()),
            "mk_spgemm_fwd");
    }
    return {out, sp_index};
}

torch::Tensor spgemm_backward(torch::Tensor ptr, torch::Tensor idx, torch::Tensor val, torch::Tensor grad_output,
                              torch::Tensor sp_index, int64_t num_nodes, int64_t num_edges, int64_t dim_sparse,
                              int64_t dim_origin) {
    check_graph(ptr, idx, val);
    cuda_contig(grad_output, "grad_output");
    cuda_contig(sp_index, "sp_index");
    TORCH_CHECK(grad_output.scalar_type() == torch::kFloat32, "grad_output must be float32");
    TORCH_CHECK(grad_output.dim() == 2, "grad_output must be 2D tensor");
    TORCH_CHECK(grad_output.size(0) == num_nodes && grad_output.size(1) == dim_origin,
                "grad_output must be [num_nodes, dim_origin]");
    TORCH_CHECK(sp_index.dim() == 2 && sp_index.size(1) == dim_sparse, "dim_sparse must equal sp_index.size(1)");
    TORCH_CHECK(dim_sparse >= 1 && dim_sparse <= dim_origin, "k must be between 1 and input dimension");
    TORCH_CHECK(ptr.numel() >= num_nodes + 1, "ptr must have num_nodes + 1 entries");
    const int ib = index_bytes(sp_index, dim_origin);
    const c10::cuda::CUDAGuard guard(grad_output.device());
    Records& rec = records_for(ptr, num_nodes);
    const int64_t n_src = sp_index.size(0);
    auto dxs = torch::empty({n_src, dim_sparse}, grad_output.options());
    py::gil_scoped_release nogil;
    chk(mk_sspmm_bwd(reinterpret_cast<const mk_part*>(rec.exec.data_ptr<int32_t>()), rec.num_parts,
                     idx.data_ptr<int32_t>(), val.data_ptr<float>(), grad_output.data_ptr<float>(), sp_index.data_ptr(),
                     ib, dxs.data_ptr<float>(), num_nodes, n_src, static_cast<int>(dim_sparse),
                     static_cast<int>(dim_origin), stream()),
        "mk_sspmm_bwd");
    return dxs;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "maxk_kernels: B200-native MaxK-GNN aggregation kernels behind the reference's extension API";
    m.def("maxk_forward", &maxk_forward, "MaxK top-k -> [N,k] values", py::arg("input"), py::arg("k"));
    m.def("maxk_forward_cbsr", &maxk_forward_cbsr, "MaxK top-k -> (sp_data, sp_index)", py::arg("input"), py::arg("k"));
    m.def("maxk_backward", &maxk_backward, "CBSR gradient -> dense", py::arg("grad_output"), py::arg("indices"));
    m.def("spgemm_forward", &spgemm_forward, "forward row-wise-product SpGEMM");
    m.def("spgemm_backward", &spgemm_backward, "backward sampled SpMM (SSpMM)");
    m.def("abi_version", []() { return mk_version(); });
}
