/*
 * maxk_b200.h -- C ABI of the B200-native MaxK-GNN aggregation hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch
 * types, no exceptions, every entry point returns 0 or a negative MK_E* code and launches
 * asynchronously on the CUDA stream it is given (a `cudaStream_t` passed as `void*`; NULL
 * is the legacy default stream).  All pointers are DEVICE pointers unless the parameter
 * name starts with `h_`.  Inputs are borrowed and never written.
 *
 * Each entry point names the reference interface it replaces.  Paths are relative to
 * julius-sk/spgemm-gnn; `so@0x...` / `.o@0x...` are addresses inside the only form the
 * reference ships its native code in (`maxk_kernels.cpython-39-x86_64-linux-gnu.so`,
 * the objects under `build/temp.linux-x86_64-cpython-39/kernels/`) as decoded in SURVEY.md section 2.3.
 *
 * CBSR ("compressed balanced sparse row"): a matrix with exactly k kept entries per row,
 *   sp_data  float32 [n, k]  row-major, the kept values,
 *   sp_index uint8   [n, k]  (index_bytes == 1, dim_origin <= 256) or
 *            uint16  [n, k]  (index_bytes == 2, dim_origin <= 65536),
 * entries of a row in ascending column order, columns of a row distinct.
 */
#ifndef MAXK_B200_H_
#define MAXK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MK_API
#else
#define MK_API __attribute__((visibility("default")))
#endif

#define MK_VERSION 202 /* major*100 + minor; 2.00: overlapped copy-engine all-gather (mk_peer_push & co. replace
                          * mk_peer_allgather / mk_peer_bank_push), mk_spgemm_fwd_banked_ex, soft peer time-outs;
                          * 2.01: mk_topk_cbsr_bank, mk_spgemm_fwd_banked_phase, mk_spgemm_fwd_banked_ln,
                          * mk_sspmm_bwd_tiled, packed tables; 2.02: NVLink multicast forms mk_peer_push_mc /
                          * mk_peer_reduce_scatter_mc.  Experimental entry points -- measured dead ends
                          * kept as the record of the experiment, off in every product path, free to go in 3.x:
                          * mk_sspmm_bwd_tma, mk_sspmm_bwd_banked, mk_spgemm_fwd_banked_phase, mk_topk_cbsr_bank. */

enum {
    MK_OK = 0,
    MK_EINVAL = -1,       /* bad argument (null pointer, k out of range, ...)            */
    MK_EUNSUPPORTED = -2, /* combination of sizes the kernels do not cover               */
    MK_ECUDA = -3,        /* a CUDA runtime call failed: see mk_last_cuda_error()        */
    MK_ENODEVICE = -4     /* no CUDA device / not an sm_100 part                         */
};

/* One record per unit of work of the forward and backward kernels: the stored entries
 * [loc, loc+len) of CSR row `row`.  Same 16-byte layout as the `.warp4` records the
 * reference reads from disk on every call (`cuda_read_array<int>`, so@0x252c0; consumed
 * with one LDG.E.128 at fwd.sass 0380).  The reference leaves the 4th word unused; here
 * it is the partial-sum slot: -1 when the row has a single record (the kernel writes the
 * output row itself), otherwise an index into the caller's partial buffer, folded in
 * fixed order by mk_spgemm_fwd (no float atomics on the forward path). */
typedef struct mk_part {
    int32_t row;
    int32_t loc;
    int32_t len;
    int32_t slot;
} mk_part;

MK_API int mk_version(void);
MK_API const char* mk_error_string(int code);
/* text of the last CUDA error this library saw on the calling thread ("" if none) */
MK_API const char* mk_last_cuda_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), else MK_ENODEVICE */
MK_API int mk_device_ok(void);

/* ---- a-1  MaxK nonlinearity -> CBSR --------------------------------------------------
 * Replaces `maxk_forward` -> `maxk_forward_cuda` -> `maxk_kernel` (maxk_cuda_kernels.o@0x1e0,
 * so@0x21110; Python caller utils/maxk_layers.py:21) and the torch formulation
 * `topk -> zeros_like -> scatter_ -> multiply` (utils/models.py:14-20).
 * Exact top-k per row: larger value first, every NaN above +inf, -0.0 == +0.0, ties go to
 * the LOWER column.  Values are bit copies of the inputs.  1 <= k <= d.                  */
MK_API int mk_topk_cbsr(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                        int index_bytes, void* stream);
/* f-3: mk_topk_cbsr and mk_cbsr_bank (below) in ONE kernel -- the row is read once, the sorted column
 * ids (what the backward and the autograd scatter need) and the banked values + cell offsets (what the
 * forward SpGEMM reads) are written once, the sorted values only if sp_data is not NULL.  Results are
 * bit-identical to the two calls.  bk_pack != NULL: the packed 8-byte form (k = 8, 16) instead of
 * bk_data / bk_slot.  Supported where mk_banked_supported(k, d) holds.                             */
MK_API int mk_topk_cbsr_bank(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                             int index_bytes, float* bk_data, uint16_t* bk_slot, void* bk_pack,
                             void* stream);

/* ---- a-2  CBSR gradient -> dense ----------------------------------------------------
 * Replaces `maxk_backward` -> `maxk_backward_cuda` (maxk_cuda_kernels.o@0x4d0: an N*k host
 * loop of `.item()` copies; Python caller utils/maxk_layers.py:40) and `grad * mask`
 * (utils/models.py:23-26).  dense[i, :] = 0; dense[i, sp_index[i,t]] = g[i,t].          */
MK_API int mk_cbsr_scatter(const float* g, const void* sp_index, int index_bytes, float* dense,
                           int64_t n, int k, int d, void* stream);

/* out[i,t] = dense[i, sp_index[i,t]] -- the vectorised replacement of the per-row Python
 * loop `_extract_sparse_format` (utils/maxk_layers.py:224-265) once the positions are
 * known, and the first half of `grad * mask`.                                           */
MK_API int mk_cbsr_gather(const float* dense, const void* sp_index, int index_bytes, float* out,
                          int64_t n, int k, int d, void* stream);

/* ---- a-5  work partition -------------------------------------------------------------
 * Replaces kernels/generate_meta.py (offline) + the per-call file read of
 * `../w12_nz64_warp_4/<graph>.warp4` (`SPMM_MAXK::do_test`, so@0x24c50-0x24d2c).
 * Cuts every CSR row into records of at most max_nz stored entries; an empty row gets one
 * record of length 0.  Records are in row order.
 *   parts == NULL : size query; *h_num_parts / *h_num_slots are written (host ints) after
 *                   the stream has been synchronised.
 *   parts != NULL : must hold the queried number of records; filled asynchronously.     */
MK_API int mk_partition(const int32_t* ptr, int64_t n_rows, int max_nz, mk_part* parts,
                        int64_t* h_num_parts, int64_t* h_num_slots, void* stream);

/* Column-blocked work list for the backward on graphs whose CBSR gradient (n_src*k*4 B) does not fit
 * in L2: mk_block_ptr cuts every CSR row (ascending column ids required) at the multiples of
 * block_width -> blk_ptr [(n_blocks+1) * n_rows]; mk_partition_ranges then builds records over the
 * n_blocks*n_rows virtual rows (row_start = blk_ptr, row_end = blk_ptr + n_rows, row_mod = n_rows,
 * skip_empty = 1), block by block, so that one launch of mk_sspmm_bwd walks the destination
 * blocks in order and its reductions stay L2-resident.  mk_partition(ptr, ...) is
 * mk_partition_ranges(ptr, ptr + 1, n, 0, max_nz, 0, ...).                                     */
MK_API int mk_block_ptr(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int n_blocks,
                        int block_width, int32_t* blk_ptr, void* stream);
MK_API int mk_partition_ranges(const int32_t* row_start, const int32_t* row_end, int64_t n_rows,
                               int64_t row_mod, int max_nz, int skip_empty, mk_part* parts,
                               int64_t* h_num_parts, int64_t* h_num_slots, void* stream);

/* ---- a-3  forward row-wise-product SpGEMM -------------------------------------------
 * Replaces `spgemm_forward` -> `spgemm_forward_cuda` -> `SPMM_MAXK::do_test` ->
 * `spmm_kernel_opt2_sparse_v3` (maxk_cuda_kernels.o@0x1260, so@0x24bf0, so@0x24b60; Python
 * callers utils/maxk_layers.py:166-171, 380-385) and DGL's `graph.update_all(copy_u, ...)`
 * (utils/models.py:163,284,407).
 *   out[r, sp_index[j,t]] += val[e] * sp_data[j,t]   for every stored e = (r <- j), t < k.
 * `out` [n_rows, d] is fully written (no zero-fill needed).  `partial` [num_slots, d] is
 * scratch, may be NULL when num_slots == 0.  Deterministic: no atomics.                  */
MK_API int mk_spgemm_fwd(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                         const int32_t* idx, const float* val, const float* sp_data,
                         const void* sp_index, int index_bytes, float* out, float* partial,
                         int64_t n_rows, int k, int d, void* stream);

/* ---- a-4  backward sampled SpMM (SSpMM) ----------------------------------------------
 * Replaces `spgemm_backward` -> `spgemm_backward_cuda` -> `SPMM_MAXK_BACKWARD::do_test` ->
 * `spmm_kernel_opt2_sparse_backward_v3` (maxk_cuda_kernels.o@0x1550, so@0x257a0) and the
 * autograd of DGL's SpMM followed by `grad * mask`.
 *   dxs[j,t] += val[e] * dy[r, sp_index[j,t]]        for every stored e = (r <- j), t < k.
 * `dxs` [n_src, k] is zero-filled by the call itself, then accumulated with vector
 * float reductions in L2 (summation order is not fixed, as in the reference's RED).      */
MK_API int mk_sspmm_bwd(const mk_part* parts, int64_t num_parts, const int32_t* idx,
                        const float* val, const float* dy, const void* sp_index, int index_bytes,
                        float* dxs, int64_t n_rows, int64_t n_src, int k, int d, void* stream);

/* mk_sspmm_bwd for CBSR gradients that do not fit L2 (products-shaped graphs: 313 MB): the sources are
 * cut into n_blocks column blocks of `block_width` (the one given to mk_block_ptr, whose output
 * blk_ptr [(n_blocks+1) * n_rows] is taken here), and the grid walks block 0 of every 8-row tile first,
 * then block 1, ...: the reductions of all CTAs in flight stay inside one L2-sized slice of dxs
 * instead of becoming DRAM read-modify-writes.  dY is staged once per (tile, block).  Same result as
 * mk_sspmm_bwd (summation order not fixed).  k in {8,16,32,64}, d % 4 == 0, ascending column ids.   */
MK_API int mk_sspmm_bwd_tiled(const int32_t* blk_ptr, int n_blocks, const int32_t* idx, const float* val,
                              const float* dy, const void* sp_index, int index_bytes, float* dxs,
                              int64_t n_rows, int64_t n_src, int k, int d, void* stream);

/* Experimental form of mk_sspmm_bwd (same contract, k == 32, uint8 ids): `tma_neighbours` (1, 2 or 4)
 * of the 4 neighbours a warp handles per step send their k contributions as ONE bulk reduction from
 * shared memory (cp.reduce.async.bulk, the TMA unit) instead of k/4 vector reductions through
 * L1TEX -- the backward is bound by the SM -> L2 request path.  Measured slower on a B200 (2.61 ms ->
 * 4.0-4.4 ms on the Reddit shape); off by default (MAXK_BWD_TMA), kept as the record of the experiment. */
MK_API int mk_sspmm_bwd_tma(const mk_part* parts, int64_t num_parts, const int32_t* idx,
                            const float* val, const float* dy, const void* sp_index, int index_bytes,
                            float* dxs, int64_t n_rows, int64_t n_src, int k, int d,
                            int tma_neighbours, void* stream);

/* ---- banked CBSR: the conflict-free internal form of the hot path ----------------------------
 * Not in the reference.  The forward / backward kernels above are bound by shared-memory bank
 * conflicts (3.6 wavefronts per access measured); mk_cbsr_bank re-orders the k entries of every
 * CBSR row and picks one of two shared-memory cells per entry so that the 8 lanes working on a
 * neighbour hit 8 different banks.  Outputs have the shapes of the inputs: bk_data / bk_index are a
 * permutation of the row (still a valid CBSR row, columns no longer ascending), bk_slot uint16
 * [n,k] is the cell offset.  The *_banked kernels compute exactly what mk_spgemm_fwd /
 * mk_sspmm_bwd compute; mk_sspmm_bwd_banked emits the gradient in the banked entry order (pair it
 * with bk_index).  Supported: k in {8,16,32,64}, d % 8 == 0, d <= 512 (mk_banked_supported).     */
MK_API int mk_banked_supported(int k, int d);
MK_API int mk_banked_rows(int d);
MK_API int mk_cbsr_bank(const float* sp_data, const void* sp_index, int index_bytes, float* bk_data,
                        void* bk_index, uint16_t* bk_slot, int64_t n, int k, int d, void* stream);
MK_API int mk_spgemm_fwd_banked(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                const int32_t* idx, const float* val, const float* bk_data,
                                const uint16_t* bk_slot, float* out, float* partial, int64_t n_rows,
                                int k, int d, void* stream);
/* The exchange a row-partitioned forward runs under (nullable everywhere it is taken).
 *   window        this rank's window of the table that is still arriving: the kernel checks done[q]
 *                 (peer.cuh) before it reads rows [q*rows_per_rank, (q+1)*rows_per_rank), and when it
 *                 has finished the whole table has arrived;
 *   h_windows     (nullable) every rank's window as mapped here, h_windows[rank] == window: the
 *                 kernel ITSELF is the all-gather -- the first `pushers` CTAs to start copy this rank's
 *                 rows of the n_seg table segments (rank r's part of segment g: h_bytes[g] bytes at
 *                 h_offsets[g] + r*h_bytes[g] of every window, multiples of 16) to the peers rank-1,
 *                 rank-2, ... over NVLink and raise done[rank] there, while the other CTAs already
 *                 compute on the blocks that have arrived.  NULL: somebody else moves the rows
 *                 (mk_peer_push, NCCL + a caller that raises the flags).                              */
typedef struct mk_fwd_exchange {
    const void* window;
    int32_t world, rank;
    int64_t rows_per_rank;
    int32_t timeout_ms; /* <= 0: 30 s */
    int32_t pushers;
    void* const* h_windows;
    int32_t n_seg;
    const int64_t* h_offsets;
    const int64_t* h_bytes;
} mk_fwd_exchange;

/* mk_spgemm_fwd_banked with the knobs of the row-partitioned and the load-balanced forms:
 *   exec_parts  (nullable) the same records in the order the CTAs should take them (maxk_kernels.py
 *               sorts them longest first, so that the grid drains on short records); `parts` stays in
 *               row order for the fold of the multi-record rows;
 *   split       (nullable) int32 [n_rows]: position in idx of the first stored entry of the row whose
 *               column is >= rank * rows_per_rank; a record walks [split, end) first, [begin, split)
 *               second -- the order in which the source blocks arrive;
 *   xchg        (nullable) see mk_fwd_exchange.                                                     */
MK_API int mk_spgemm_fwd_banked_ex(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                   const mk_part* exec_parts, const int32_t* idx, const float* val,
                                   const float* bk_data, const uint16_t* bk_slot, float* out,
                                   float* partial, int64_t n_rows, int k, int d, const int32_t* split,
                                   const mk_fwd_exchange* xchg, void* stream);
/* One phase of a forward that is cut by SOURCE BLOCK ("chunk the all-gather by
 * source rank and start on the local block first").  blk_ptr is mk_block_ptr's output for n_blocks
 * blocks: blk_ptr[b * row_stride + r] = position in idx of row r's first entry of block b
 * (b = n_blocks: the row's end).  The launch walks the entries of blocks [a0, a1), then [b0, b1) (two
 * ranges, because "rank, rank+1, ..." wraps around); `accumulate` adds to the rows the earlier phases
 * wrote instead of overwriting them; `last` keeps the completion contract of mk_fwd_exchange (the kernel
 * returns when the whole table has arrived) -- earlier phases only wait for the blocks they read.
 * The phases of one forward must cover every block exactly once; the row sums then differ from the
 * one-launch forward only by the association of the per-phase partial sums.                         */
typedef struct mk_fwd_phase {
    const int32_t* blk_ptr;
    int64_t row_stride;
    int32_t n_blocks;
    int32_t a0, a1, b0, b1;
    int32_t accumulate;
    int32_t last;
} mk_fwd_phase;
MK_API int mk_spgemm_fwd_banked_phase(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                      const mk_part* exec_parts, const int32_t* idx, const float* val,
                                      const float* bk_data, const uint16_t* bk_slot, float* out,
                                      float* partial, int64_t n_rows, int k, int d,
                                      const mk_fwd_exchange* xchg, const mk_fwd_phase* phase,
                                      void* stream);
/* f-3: the forward SpGEMM with the layer's epilogue applied to every row as it is finished:
 *     y[r] = LayerNorm(h_self[r] + (A x Xs)[r] + bias) * gamma + beta
 * -- `output = h_self + aggregated_feat; output = self.norm(output)` (utils/maxk_layers.py:174-182)
 * without writing the aggregated row and reading it back.  h_self and bias may be NULL; z (the
 * pre-normalisation sum), mean, rstd are what mk_layernorm_bwd needs and may be NULL for inference.
 * Bit-identical to mk_spgemm_fwd_banked / _packed followed by mk_add_layernorm_fwd.  `table` is
 * bk_data with bk_slot, or the packed table with bk_slot == NULL.  d % 4 == 0, d <= 512; 16-byte
 * aligned pointers.                                                                                 */
typedef struct mk_fwd_epilogue {
    const float* h_self;
    const float* bias;
    const float* gamma;
    const float* beta;
    float* z;
    float* mean;
    float* rstd;
    float eps;
} mk_fwd_epilogue;
MK_API int mk_spgemm_fwd_banked_ln(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                   const mk_part* exec_parts, const int32_t* idx, const float* val,
                                   const void* table, const uint16_t* bk_slot, float* y, float* partial,
                                   int64_t n_rows, int k, int d, const mk_fwd_epilogue* epilogue,
                                   void* stream);
/* Packed banked CBSR for k = 8, 16 (mk_packed_supported): one entry = {float value, uint16 cell,
 * uint16 column} in 8 bytes, bk_pack [n, k] of them, so that a lane fetches value and cell offset with
 * ONE load -- at these widths the forward is bound by L1 wavefronts per gathered row, and the separate
 * value / offset arrays cost two.  Same result as mk_spgemm_fwd; arguments as mk_spgemm_fwd_banked_ex. */
MK_API int mk_packed_supported(int k, int d);
MK_API int mk_cbsr_bank_packed(const float* sp_data, const void* sp_index, int index_bytes,
                               void* bk_pack, int64_t n, int k, int d, void* stream);
MK_API int mk_spgemm_fwd_packed_ex(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                   const mk_part* exec_parts, const int32_t* idx, const float* val,
                                   const void* bk_pack, float* out, float* partial, int64_t n_rows,
                                   int k, int d, const int32_t* split, const mk_fwd_exchange* xchg,
                                   void* stream);
MK_API int mk_sspmm_bwd_banked(const mk_part* parts, int64_t num_parts, const int32_t* idx,
                               const float* val, const float* dy, const uint16_t* bk_slot,
                               float* dxs, int64_t n_rows, int64_t n_src, int k, int d, void* stream);

/* ---- f-3  epilogue of the aggregation: z = a + b + bias, y = LayerNorm(z) * gamma + beta -------
 * Replaces `output = h_self + aggregated_feat; output = self.norm(output)`
 * (utils/maxk_layers.py:174-182; nn.LayerNorm in utils/models.py:122, 260, 382) and its autograd.
 * b, bias, z may be NULL.  d % 4 == 0, d <= 1024, all pointers 16-byte aligned.
 * Backward: gz = d loss / d z (== d/da == d/db), dgamma, dbeta, dbias (= column sums of gz, may be
 * NULL); `workspace` holds 3 * mk_layernorm_parts() * d floats; the parameter gradients are folded
 * in fixed order.                                                                               */
MK_API int mk_layernorm_parts(void);
MK_API int mk_add_layernorm_fwd(const float* a, const float* b, const float* bias,
                                const float* gamma, const float* beta, float* z, float* y,
                                float* mean, float* rstd, int64_t n, int d, float eps, void* stream);
MK_API int mk_layernorm_bwd(const float* gy, const float* z, const float* gamma, const float* mean,
                            const float* rstd, float* gz, float* dgamma, float* dbeta, float* dbias,
                            float* workspace, int64_t n, int d, void* stream);

/* ---- e  peer-memory exchange of the row-partitioned path (multi-GPU, one process per GPU) ------
 * Not in the reference (single GPU; README_INTEGRATED.md:382 lists "Multi-GPU support with NCCL" as
 * future work).  SURVEY.md section 8e: 1-D row partition, all-gather of the CBSR table in front of
 * the forward SpGEMM, reduce-scatter of the CBSR gradient behind the backward SSpMM.  These entry
 * points are that pair over peer-mapped memory (NVLink), next to the torch.distributed/NCCL form in
 * spgemm_gnn_b200/dist.py.
 *
 * A WINDOW is one device buffer per rank (same size on every rank) that every peer process maps
 * with CUDA IPC.  Its first MK_PEER_HEADER_BYTES are flags owned by the library (zeroed by
 * mk_peer_alloc, never to be written by the caller); payload offsets below count from the window
 * base and must be >= MK_PEER_HEADER_BYTES and 16-byte aligned.  `h_windows[q]` is rank q's window
 * in the caller's address space (own window: the mk_peer_alloc pointer; others: mk_peer_open).
 * All ranks must issue the same collectives on a window in the same order.  A kernel that waits for
 * a peer longer than `timeout_ms` (<= 0: 30 s) writes the collective's number into the header's
 * error word and stops waiting (its results are then undefined; mk_peer_epoch reports the word).
 *
 * Overlapped all-gather (the forward).  Per collective, on the rank's main stream unless noted:
 *   mk_peer_begin_push   one-block kernel: returns when every peer has released the table buffer
 *                        (0 or 1) that is about to be overwritten;
 *   <producer>           the rank's rows are written into its OWN window (mk_cbsr_bank outputs, ...);
 *   mk_peer_publish      opens collective e = epoch + 1: the own rows are complete;
 *   mk_peer_push         on a SIDE stream, after an event recorded behind mk_peer_publish: one
 *                        cudaMemcpyAsync per peer and segment (copy engines, no SM), peers visited
 *                        rank-1, rank-2, ...; each peer's copies are followed by a 4-byte copy of e
 *                        into done[rank] of that peer's header.  Segment g of this rank is the
 *                        h_bytes[g] bytes at h_offsets[g] + rank * h_bytes[g] of every window;
 *   <consumer>           mk_spgemm_fwd_banked_ex(wait_window = own window) starts at once and waits
 *                        per source block; any other consumer calls mk_peer_wait_all first;
 *   mk_peer_release      after the last reader of the table: tells the peers this rank is done with it. */
#define MK_PEER_MAX_RANKS 16
#define MK_PEER_HEADER_BYTES 1024
#define MK_PEER_HANDLE_BYTES 64
MK_API int mk_peer_alloc(int64_t bytes, void** window);
MK_API int mk_peer_free(void* window);
MK_API int mk_peer_export(void* window, unsigned char* h_handle /* [MK_PEER_HANDLE_BYTES] */);
MK_API int mk_peer_open(const unsigned char* h_handle, void** window);
MK_API int mk_peer_close(void* window);
/* collectives completed through the window and its error word (synchronises the stream)          */
MK_API int mk_peer_epoch(const void* window, uint32_t* h_epoch, uint32_t* h_error, void* stream);
MK_API int mk_peer_begin_push(void* window, int world, int rank, int buffer, int timeout_ms, void* stream);
MK_API int mk_peer_publish(void* window, int rank, int buffer, void* stream);
MK_API int mk_peer_push(void* const* h_windows, int world, int rank, int n_seg,
                        const int64_t* h_offsets, const int64_t* h_bytes, void* stream);
/* the steps first_step, first_step + step_stride, ... of mk_peer_push (step s sends to rank - s):
 * several side streams, one call each, put several copy engines to work while every stream still
 * visits its peers nearest first                                                                   */
MK_API int mk_peer_push_steps(void* const* h_windows, int world, int rank, int n_seg,
                              const int64_t* h_offsets, const int64_t* h_bytes, int first_step,
                              int step_stride, void* stream);
/* the same all-gather by NVLink stores from `pushers` CTAs of 32 threads, as a kernel of its own (what
 * the forward kernel's pusher CTAs do, mk_fwd_exchange): after mk_peer_publish, on the main stream    */
MK_API int mk_peer_push_sm(void* const* h_windows, int world, int rank, int n_seg,
                           const int64_t* h_offsets, const int64_t* h_bytes, int pushers, void* stream);
MK_API int mk_peer_wait_all(void* window, int world, int timeout_ms, void* stream);
MK_API int mk_peer_release(void* const* h_windows, int world, int rank, void* stream);
/* Reduce-scatter by loads (the backward): out[0 .. block_bytes/4) = sum over q (rank order, fixed) of
 * the floats at offset + rank * block_bytes of rank q's window.  When the call's kernel has
 * finished, no peer reads this rank's window any more (it may be overwritten).  `grid` is the number
 * of thread blocks (0: the library's choice); it is capped so that the grid is resident at once.
 * The _virtual form runs ALL ranks of a single-process emulation (every window on one device) in
 * one launch -- separate launches that wait on one another must not share a device; tests only.   */
MK_API int mk_peer_reduce_scatter(void* const* h_windows, int world, int rank, int64_t offset,
                                  int64_t block_bytes, float* out, int grid, int timeout_ms,
                                  void* stream);
/* NVLink multicast (NVLS) forms of the two exchanges, for windows that are symmetric memory bound to a
 * multicast object (peer.py allocates them through torch's symmetric-memory allocator): `mc_window` is
 * the multicast address of the window -- a store to it lands at the same offset of every rank's
 * window, a multimem.ld_reduce from it returns the sum over all of them.
 *   mk_peer_push_mc            the all-gather of mk_peer_push_sm with every row sent ONCE (after
 *                              mk_peer_publish; raises done[rank] in every peer's header at the end);
 *   mk_peer_reduce_scatter_mc  mk_peer_reduce_scatter with the sum formed by the switch (one reduced
 *                              16-byte load per element instead of `world` loads; the order of the sum
 *                              is the switch's, as with NCCL).                                        */
MK_API int mk_peer_push_mc(void* const* h_windows, void* mc_window, int world, int rank, int n_seg,
                           const int64_t* h_offsets, const int64_t* h_bytes, int grid, void* stream);
MK_API int mk_peer_reduce_scatter_mc(void* const* h_windows, const void* mc_window, int world, int rank,
                                     int64_t offset, int64_t block_bytes, float* out, int grid,
                                     int timeout_ms, void* stream);
MK_API int mk_peer_reduce_scatter_virtual(void* const* h_windows, int world, int64_t offset,
                                          int64_t block_bytes, float* const* h_outs, int grid,
                                          int timeout_ms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAXK_B200_H_ */
