// launch.hpp -- host-callable launchers implemented in kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "device_types.hpp"

namespace ikc {

// Generic two-launch path (vertical -> f32 tmp in HBM -> horizontal).  `job` holds device pointers.
cudaError_t launch_generic(const DevJob& job, bool exact, cudaStream_t stream);

// Fused ring kernel over a list of work items (device arrays `jobs`, `items`).
bool fused_supported(int channels, int ring_k_v, int ring_k_h);
int fused_max_src_bytes(int channels);   // source bytes of one strip row the kernel can stage
int fused_group_rows();                  // intermediate rows per group
size_t fused_smem_bytes(int channels, int ring_k_v, int ring_k_h, const FusedGeom& geom);
// words_per_thread: 1 = 8 warps per CTA (one source word per thread per row), 2 = 4 warps per CTA
cudaError_t launch_fused(int channels, int ring_k_v, int ring_k_h, int words_per_thread, const DevJob* jobs, const WorkItem* items,
                         const FusedGeom& geom, cudaStream_t stream);

// Tile kernel (tile.cu): output-stationary fused passes over shared-memory tiles.
size_t tile_smem_bytes(const TileGeom& geom);
cudaError_t launch_tile(const DevJob* jobs, const WorkItem* items, const TileGeom& geom, cudaStream_t stream);

}  // namespace ikc
