// Host helper: encode a 3-D bf16 tensor map [batches][rows][inner] with 128-byte swizzle and zero
// out-of-bounds fill (cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint, so the
// library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>

namespace fs2 {
int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, long long inner, long long rows, long long batches,
                      long long ld, long long batch_stride, int box_inner, int box_rows);
}
