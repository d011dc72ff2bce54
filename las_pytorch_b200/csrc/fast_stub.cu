// Placeholder for the parts of the LAS_MODE_BF16 path whose kernels are still being brought up: fails loudly.
#include "las_fast.cuh"
namespace las {
bool fast_available() { return false; }
static int nyi() { return fail(LAS_EINVAL, "LAS_MODE_BF16 speller is not built into this library yet"); }
size_t fast_speller_packed_bytes(const las_speller_dims*) { return 256; }
int fast_speller_pack(const las_speller_weights*, const las_speller_dims*, void*, cudaStream_t) { return nyi(); }
size_t fast_speller_workspace_bytes(const las_speller_dims*, int) { return 256; }
int fast_speller_decode(const las_decode_io*, const void*, const void*, const las_speller_dims*, int, int, int, void*, void*,
                        cudaStream_t) { return nyi(); }
}  // namespace las
