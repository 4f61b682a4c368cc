// fir_fast.cu -- fast convolution kernel (placeholder until the tiled kernel lands).
#include "fir_kernels.h"

namespace rsb {
bool fast_supported(uint32_t, uint32_t, double) { return false; }
uint32_t fast_tile_out(uint32_t, uint32_t, double) { return kExactTileOut; }
uint32_t fast_streams_per_group(uint32_t, uint32_t, double) { return kExactStreamsPerGroup; }
void launch_conv_fast(const ConvParams &, double, uint32_t, int, cudaStream_t) {}
}  // namespace rsb
