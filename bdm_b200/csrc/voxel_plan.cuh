// voxel_plan.cuh -- layout of the per-shape voxel plan that bdm_voxel_plan leaves in the caller's workspace
// (written by vox_sort_kernel in voxelize.cu; read by the dense / compact fills and by sparse_conv.cu).
#pragma once
#include "common.cuh"

namespace bdm {

constexpr int kFastMaxR3 = 32768;
constexpr int kFastMaxN = 16384;

struct VoxAuxLayout {
  size_t header;   // u32[4]: [0] = number of occupied voxels
  size_t bitmask;  // u32[nw]
  size_t obase;    // u16[nw]
  size_t ostart;   // u16[n+1]
  size_t rank;     // u16[n]
  size_t stride;   // bytes per shape
  int nw;
};

__host__ __device__ inline VoxAuxLayout vox_aux_layout(int n, int r3) {
  VoxAuxLayout L;
  L.nw = (r3 + 31) / 32;
  size_t off = 0;
  L.header = off;  off += 16;
  L.bitmask = off; off = align_up(off + sizeof(uint32_t) * L.nw, 16);
  L.obase = off;   off = align_up(off + sizeof(uint16_t) * L.nw, 16);
  L.ostart = off;  off = align_up(off + sizeof(uint16_t) * (n + 1), 16);
  L.rank = off;    off = align_up(off + sizeof(uint16_t) * n, 16);
  L.stride = off;
  return L;
}

inline bool vox_fast_path(int n, int r3) { return r3 <= kFastMaxR3 && n <= kFastMaxN && n >= 1; }

inline int check_workspace(const VoxAuxLayout &L, int b, const void *workspace, size_t workspace_bytes) {
  if (workspace == nullptr) return BDM_ERR_NULL_POINTER;
  if (workspace_bytes < L.stride * (size_t)b) return BDM_ERR_WORKSPACE_TOO_SMALL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return BDM_ERR_MISALIGNED;
  return BDM_OK;
}

}  // namespace bdm
