// Miniatures of the node-assignment chain on ONE warp: cycles per placement for
//   A  64-bit guard key, two redux.sync           (chain_swar, general)
//   B  32-bit compact compare key, one redux.sync (chain_swar<K32>)
// with the rare paths behind unlikely branches, records prefetched two ahead, two placements per trip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/chain tools/ubench/chain.cu && tools/ubench/chain
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
#define NREC 2048
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
  unsigned hi = (unsigned)(k >> 32);
  unsigned mhi = __reduce_min_sync(FULL, hi);
  unsigned lo = (hi == mhi) ? (unsigned)k : 0xFFFFFFFFu;
  unsigned mlo = __reduce_min_sync(FULL, lo);
  return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __noinline__ void rare(unsigned* have, unsigned long long* Y, unsigned s) {
  *have |= 1u << s;
  *Y += 1ull << 40;
}
template <int MODE>
__global__ void k_chain(long long* out, uint4* grec, unsigned seed, int reps) {
  __shared__ uint4 rec[NREC];
  const unsigned l = threadIdx.x;
  for (unsigned i = l; i < NREC; i += 32) rec[i] = grec[i];
  __syncwarp();
  unsigned long long G = 0x0000220020000000ull << 1;  // guard bits of three fields
  asm volatile("mov.b64 %0, %0;" : "+l"(G));
  unsigned long long Y = (((unsigned long long)(seed + l * 977u) & 0x7ffffffull) << 18 | l | (G >> 1)) << 1;
  unsigned ck = ((seed + l * 977u) & 0x3ffffffu) << 6 | l << 1;
  unsigned mystat = 0xffffffffu, fullstat = 0xffffffffu, have = 0xffffffffu, acc = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    const unsigned total = NREC, last = NREC - 1;
    auto step = [&](const uint4 rc, unsigned tt) -> bool {
      const unsigned long long preq = ((unsigned long long)rc.y << 32) | rc.x;
      const unsigned s = rc.z & 0xFFu;
      if (__builtin_expect(!((have >> s) & 1u), 0)) rare(&have, &Y, s);
      const unsigned long long d = Y - preq;
      const bool fit = ((d & G) == G) && ((mystat >> s) & 1u);
      if (MODE == 0) {
        const unsigned long long m = warp_min_u64(fit ? Y : ~0ull);
        if (__builtin_expect(m == ~0ull, 0)) return true;
        if (Y == m) {
          Y = d & ~1ull;
          mystat = fullstat;
        }
        if (l == 0) rec[tt].w = (unsigned)(m >> 1) & 0x1ffffu;
        if (__builtin_expect((unsigned)m & 1u, 0)) rare(&have, &Y, s);
      } else {
        const unsigned m = __reduce_min_sync(FULL, fit ? ck : 0xFFFFFFFFu);
        if (__builtin_expect(m == 0xFFFFFFFFu, 0)) return true;
        if (ck == m) {
          Y = d;
          ck -= rc.w;
          mystat = fullstat;
          rec[tt].w = l;
        }
        if (__builtin_expect(m & 1u, 0)) rare(&have, &Y, s);
      }
      return false;
    };
    unsigned t = 0;
    uint4 ra = rec[0], rb = rec[1];
    for (; t + 1 < total; t += 2) {
      if (step(ra, t)) break;
      ra = rec[t + 2 <= last ? t + 2 : last];
      if (step(rb, t + 1)) break;
      rb = rec[t + 3 <= last ? t + 3 : last];
    }
    acc += t;
    // top the rows up again so that the next repetition fits too
    Y = (((unsigned long long)(seed + l * 977u + r) & 0x7ffffffull) << 18 | l | (G >> 1)) << 1;
    ck = ((seed + l * 977u + r) & 0x3ffffffu) << 6 | l << 1;
  }
  long long t1 = clock64();
  if (l == 0) {
    out[0] = t1 - t0;
    out[1] = acc;
  }
  if (Y == 1 && ck == 7 && have == 3) out[2] = 1;
}
int main() {
  long long *d, h[3];
  uint4 *rec, hr[NREC];
  for (int i = 0; i < NREC; ++i) {
    // small requests in each field so that most lanes fit; window in z; compact request in w
    unsigned long long pq = ((unsigned long long)(1 + i % 3) << 18 | (unsigned long long)(i % 5) << 29 | (unsigned long long)(i % 2) << 44) << 1;
    hr[i] = make_uint4((unsigned)pq, (unsigned)(pq >> 32), (unsigned)(i % 5), (1u + i % 3) << 6);
  }
  cudaMalloc(&d, sizeof(h));
  cudaMalloc(&rec, sizeof(hr));
  cudaMemcpy(rec, hr, sizeof(hr), cudaMemcpyHostToDevice);
  const int reps = 8;
  for (int mode = 0; mode < 2; ++mode) {
    for (int w = 0; w < 2; ++w) {
      if (mode == 0) k_chain<0><<<1, 32>>>(d, rec, 12345u, reps);
      else k_chain<1><<<1, 32>>>(d, rec, 12345u, reps);
      cudaMemcpy(rec, hr, sizeof(hr), cudaMemcpyHostToDevice);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s: %.1f cycles per placement (%lld placements)\n", mode == 0 ? "A 64-bit key, 2x redux" : "B 32-bit key, 1x redux", (double)h[0] / (double)h[1], h[1]);
  }
  return 0;
}
