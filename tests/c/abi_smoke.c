/* The C ABI from plain C (no Python, no C++, no torch): include/maxk_b200.h must compile as C99,
 * the library must link, and argument validation must answer before any CUDA call, so this runs
 * on a box without a GPU.  Built and run by tests/test_abi.py::test_header_is_plain_c_and_links. */
#include <stdio.h>
#include <string.h>

#include "maxk_b200.h"

#define CHECK(cond)                                                   \
    do {                                                              \
        if (!(cond)) {                                                \
            fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond); \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main(void) {
    mk_part rec;
    void* windows[MK_PEER_MAX_RANKS] = {0};
    unsigned char handle[MK_PEER_HANDLE_BYTES];
    void* p = 0;

    CHECK(sizeof(rec) == 16);                       /* the .warp4 record layout */
    CHECK(mk_version() == MK_VERSION);
    CHECK(strcmp(mk_error_string(MK_OK), "ok") == 0);
    CHECK(strcmp(mk_error_string(MK_EINVAL), "invalid argument") == 0);
    CHECK(mk_last_cuda_error() != 0);

    /* every entry point rejects nonsense before touching the device */
    CHECK(mk_topk_cbsr(0, 4, 16, 0, 0, 0, 1, 0) == MK_EINVAL);
    CHECK(mk_topk_cbsr(0, 0, 16, 8, 0, 0, 1, 0) == MK_OK);
    CHECK(mk_cbsr_scatter(0, 0, 3, 0, 4, 8, 16, 0) == MK_EINVAL);
    CHECK(mk_partition(0, -1, 64, 0, 0, 0, 0) == MK_EINVAL);
    CHECK(mk_spgemm_fwd(0, 5, 0, 0, 0, 0, 0, 1, 0, 0, 5, 8, 4, 0) == MK_EINVAL);
    CHECK(mk_sspmm_bwd(0, 1, 0, 0, 0, 0, 7, 0, 1, 1, 8, 16, 0) == MK_EINVAL);
    CHECK(mk_banked_supported(32, 256) == 1 && mk_banked_supported(7, 256) == 0);
    CHECK(mk_peer_alloc(8, &p) == MK_EINVAL);
    CHECK(mk_peer_export(0, handle) == MK_EINVAL);
    CHECK(mk_peer_push(windows, 2, 0, 1, 0, 0, 0) == MK_EINVAL);
    CHECK(mk_peer_publish(0, 0, 0, 0) == MK_EINVAL);
    CHECK(mk_peer_reduce_scatter(windows, 2, 5, 1024, 64, 0, 0, 0, 0) == MK_EINVAL);
    printf("c abi ok: version %d\n", mk_version());
    return 0;
}
