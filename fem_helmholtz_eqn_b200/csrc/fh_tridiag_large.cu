// placeholder until the partitioned large-system solver lands
#include "fh_common.cuh"
namespace fh {
size_t solve_large_workspace_bytes(int64_t) { return 0; }
int solve_large(const cplx *, const cplx *, const cplx *, const cplx *, cplx *, int64_t, int32_t *, void *, size_t,
                cudaStream_t) {
    set_error("large-system solver not built yet");
    return FH_E_UNSUPPORTED;
}
}  // namespace fh
extern "C" {
size_t fh_tridiag_solve_large_workspace_bytes(int64_t n) { return fh::solve_large_workspace_bytes(n); }
int fh_tridiag_solve_large_c128(const fh_c128 *, const fh_c128 *, const fh_c128 *, const fh_c128 *, fh_c128 *, int64_t,
                                int32_t *, void *, size_t, fh_stream_t) {
    fh::set_error("large-system solver not built yet");
    return FH_E_UNSUPPORTED;
}
}
