// launch.hpp -- host-callable launchers of the CUDA kernels (generic.cu, fused.cu / fused_conv.cu, tile.cu, up2.cu).
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
// True when the kernel has loops specialised for these uniform steps (PassPlan::uni_step of both passes).
bool fused_has_uniform(int channels, int ring_k_v, int ring_k_h, int step_v, int step_h);
// step_v / step_h: uniform steps the launch is specialised for (both 0: general loops only).
// convert: the jobs of the launch store a different channel count than they read (DevJob::out_channels).
cudaError_t launch_fused(int channels, int ring_k_v, int ring_k_h, int step_v, int step_h, bool convert,
                         const DevJob* jobs, const WorkItem* items, const FusedGeom& geom, cudaStream_t stream);

// Banded kernel (banded.cu / banded_conv.cu): downscales with the vertical pass on the tensor cores.
bool banded_supported(int channels, int band_n);
int banded_max_src_bytes();              // source bytes of one strip row the kernel stages
int banded_group_rows();                 // intermediate rows per group
size_t banded_smem_bytes(int channels, const BandGeom& geom);
size_t banded_max_smem();                // the kernel's shared-memory budget (two CTAs per SM)
cudaError_t launch_banded(int channels, bool convert, const DevJob* jobs, const WorkItem* items, const BandGeom& geom,
                          cudaStream_t stream);

// Banded8 kernel (banded8.cu / banded8_conv.cu): downscales with the vertical pass as an integer product on the tensor
// cores, the source bytes used as they are (no conversion pass).
bool banded8_supported(int channels, int limbs);
int banded8_max_src_bytes();
int banded8_tile_rows();                 // intermediate rows per horizontal phase: row chunks are cut on multiples of it
size_t banded8_smem_bytes(int channels, const Band8Geom& geom);
size_t banded8_max_smem();
cudaError_t launch_banded8(int channels, bool convert, const DevJob* jobs, const WorkItem* items, const Band8Geom& geom,
                           cudaStream_t stream);

// Banded8t kernel (banded8t.cu): Rgba8 downscales that are exactly 2:1 horizontally; accumulator lanes are output rows and
// the horizontal pass runs from registers.  Work items are (band of the pass's band8t_rows output rows) x (column range).
size_t banded8t_smem_bytes(const Band8TGeom& geom);
cudaError_t launch_banded8t(const DevJob* jobs, const WorkItem* items, const Band8TGeom& geom, cudaStream_t stream);

// Banded8u kernel (banded8u.cu): exact 2x upscales of Rgb8 / Rgba8 with the vertical pass on the tensor cores (accumulator
// lanes are output rows) and the horizontal pass from registers.  Work items: (band of 128 output rows) x (column range
// in whole blocks of 64 output pixels).
bool banded8u_supported(int channels, int taps_h, int off_h, int chunks_v);
int banded8u_band_rows();
cudaError_t launch_banded8u(int channels, int taps, const DevJob* jobs, const WorkItem* items, int n_items, cudaStream_t stream);

// Tile kernel (tile.cu): output-stationary fused passes over shared-memory tiles.
size_t tile_smem_bytes(const TileGeom& geom);
cudaError_t launch_tile(int bytes_per_sample, const DevJob* jobs, const WorkItem* items, const TileGeom& geom,
                        cudaStream_t stream);

// Exact 2x upscale kernel (up2.cu).  Work items are output tiles of up2_tile_w(channels) x up2_tile_h().
bool up2_supported(int channels, int taps_v, int taps_h);
int up2_tile_w(int channels);
int up2_tile_h();
cudaError_t launch_up2(int channels, int taps, const DevJob* jobs, const WorkItem* items, int n_items, cudaStream_t stream);

}  // namespace ikc
