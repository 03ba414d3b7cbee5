// Host-side check of brief_cube_voxel (brief_common.cuh) — the __host__ __device__ function gen_indices_kernel calls —
// so that the window -> voxel arithmetic can be pinned against the oracle without a GPU (tests/test_cubes.py).
//   cube_index_host h w ch cw cube_vox cube_id...   -> one line per cube: its cube_vox voxel indices
#include <cstdio>
#include <cstdlib>

#include "../../brief_pytorch_b200/csrc/brief_common.cuh"

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int h = atoi(argv[1]), w = atoi(argv[2]), ch = atoi(argv[3]), cw = atoi(argv[4]);
  const long long vox = atoll(argv[5]);
  for (int a = 6; a < argc; ++a) {
    const long long cube = atoll(argv[a]);
    for (long long o = 0; o < vox; ++o) printf(o ? " %lld" : "%lld", brief_cube_voxel(h, w, ch, cw, cube, o));
    printf("\n");
  }
  return 0;
}
