// Placeholder for the LAS_MODE_BF16 path while its kernels are being brought up: every entry fails loudly.
#include "las_fast.cuh"
namespace las {
bool fast_available() { return false; }
static int nyi() { return fail(LAS_EINVAL, "LAS_MODE_BF16 is not built into this library yet"); }
size_t fast_listener_packed_bytes(const las_listener_dims*) { return 256; }
int fast_listener_pack(const las_lstm_weights*, const las_listener_dims*, void*, cudaStream_t) { return nyi(); }
size_t fast_listener_workspace_bytes(const las_listener_dims*) { return 256; }
int fast_listener_forward(const float*, const void*, const las_listener_dims*, float*, void*, cudaStream_t) { return nyi(); }
size_t fast_speller_packed_bytes(const las_speller_dims*) { return 256; }
int fast_speller_pack(const las_speller_weights*, const las_speller_dims*, void*, cudaStream_t) { return nyi(); }
size_t fast_speller_workspace_bytes(const las_speller_dims*, int) { return 256; }
int fast_speller_decode(const las_decode_io*, const void*, const void*, const las_speller_dims*, int, int, int, void*, void*,
                        cudaStream_t) { return nyi(); }
}  // namespace las
