// pxb_sort.cuh -- device-wide LSD radix sort (u64 key, u32 payload) and exclusive scan with the element
// count read from DEVICE memory, so no host synchronisation is needed between pipeline stages.
//
// Replaces the role of the reference's radix sort (physx/source/gpucommon/src/CUDA/RadixSort.cuh,
// radixSortImpl.cu:37-237: 4 bits/pass, 2 kernels/pass, fixed grid) in the broadphase.
// Design: 8 bits per pass, fixed persistent grid sized to the SM count, each CTA owns one contiguous
// chunk per pass (histogram -> per-digit row scan of the [CTA][256] counts -> stable scatter using
// warp-level __match_any_sync ranking).  Only ceil(keyBits/8) passes run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_MAX_CTAS 592  // 148 SMs x 4

__device__ __forceinline__ void rs_chunk(uint32_t n, uint32_t& begin, uint32_t& end) {
  const uint32_t G = gridDim.x;
  uint32_t chunk = (n + G - 1) / G;
  chunk = (chunk + RS_THREADS - 1) / RS_THREADS * RS_THREADS;
  begin = min(n, blockIdx.x * chunk);
  end = min(n, begin + chunk);
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ dN, uint32_t shift, uint32_t* __restrict__ blockHist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint32_t b, e; rs_chunk(*dN, b, e);
  for (uint32_t i = b + threadIdx.x; i < e; i += RS_THREADS) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 0xffu], 1u);
  __syncthreads();
  blockHist[blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];
}

// One CTA per digit: exclusive scan of that digit's counts over the CTAs (in place) + the digit total.
// The cross-digit prefix (256 values) is rebuilt by every scatter CTA in shared memory, so a pass is
// hist -> row scan (256 tiny CTAs) -> scatter, with no serial single-CTA step.
__global__ void __launch_bounds__(128) k_rs_scan(uint32_t* __restrict__ blockHist, uint32_t G, uint32_t* __restrict__ digitTotals) {
  __shared__ uint32_t ws[4];
  __shared__ uint32_t carry;
  const uint32_t d = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < G; base += 128) {
    const uint32_t g = base + t;
    const uint32_t v = g < G ? blockHist[g * 256 + d] : 0;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { const uint32_t c = ws[w]; if (w < (int)warp) woff += c; tot += c; }
    if (g < G) blockHist[g * 256 + d] = carry + woff + x - v;
    __syncthreads();
    if (t == 0) carry += tot;
    __syncthreads();
  }
  if (t == 0) digitTotals[d] = carry;
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                           uint64_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                           const uint32_t* __restrict__ dN, uint32_t shift, const uint32_t* __restrict__ blockHist,
                                                           const uint32_t* __restrict__ digitTotals) {
  __shared__ uint32_t running[256];
  __shared__ uint32_t warpCnt[RS_WARPS][256];
  __shared__ uint32_t wsum[RS_WARPS];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {  // exclusive prefix of the 256 digit totals (block-wide scan), plus this CTA's offset inside the digit
    const uint32_t v = digitTotals[tid];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) if (w < (int)warp) woff += wsum[w];
    running[tid] = woff + x - v + blockHist[blockIdx.x * 256 + tid];
  }
  uint32_t b, e; rs_chunk(*dN, b, e);
  for (uint32_t base = b; base < e; base += RS_THREADS) {
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) warpCnt[w][tid] = 0;
    __syncthreads();
    const uint32_t i = base + tid;
    const bool valid = i < e;
    uint64_t key = 0; uint32_t val = 0; uint32_t digit = 0x100u + lane;  // invalid lanes never match a real digit
    if (valid) { key = keysIn[i]; val = valsIn[i]; digit = (uint32_t)(key >> shift) & 0xffu; }
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warpCnt[warp][digit] = __popc(peers);
    __syncthreads();
    {  // thread `tid` owns digit `tid`: turn per-warp counts into offsets and advance the running base
      uint32_t off = running[tid];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) { const uint32_t c = warpCnt[w][tid]; warpCnt[w][tid] = off; off += c; }
      running[tid] = off;
    }
    __syncthreads();
    if (valid) { const uint32_t pos = warpCnt[warp][digit] + rank; keysOut[pos] = key; valsOut[pos] = val; }
    __syncthreads();
  }
}

struct RadixSortTemp { uint32_t* blockHist; uint32_t* digitTotals; uint32_t ctas; };

// Sorts (keys, vals)[0..*dN) on `stream`. Returns 0 if the result is in (keys, vals), 1 if in (keysAlt, valsAlt).
static inline int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keysAlt, uint32_t* valsAlt, const uint32_t* dN,
                                   uint32_t keyBits, const RadixSortTemp& tmp, cudaStream_t stream) {
  const uint32_t passes = (keyBits + 7) / 8;
  int cur = 0;
  for (uint32_t p = 0; p < passes; ++p) {
    const uint64_t* kin = cur ? keysAlt : keys; const uint32_t* vin = cur ? valsAlt : vals;
    uint64_t* kout = cur ? keys : keysAlt; uint32_t* vout = cur ? vals : valsAlt;
    k_rs_hist<<<tmp.ctas, RS_THREADS, 0, stream>>>(kin, dN, p * 8, tmp.blockHist);
    k_rs_scan<<<256, 128, 0, stream>>>(tmp.blockHist, tmp.ctas, tmp.digitTotals);
    k_rs_scatter<<<tmp.ctas, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, dN, p * 8, tmp.blockHist, tmp.digitTotals);
    cur ^= 1;
  }
  return cur;
}

// ---- exclusive scan of u32 with device-side count (same chunking) ----
__global__ void __launch_bounds__(RS_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, const uint32_t* __restrict__ dN, uint32_t* __restrict__ blockSums) {
  __shared__ uint32_t ws[RS_WARPS];
  uint32_t b, e; rs_chunk(*dN, b, e);
  uint32_t s = 0;
  for (uint32_t i = b + threadIdx.x; i < e; i += RS_THREADS) s += in[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < RS_WARPS; ++w) t += ws[w]; blockSums[blockIdx.x] = t; }
}
// single CTA; also writes the grand total to *dTotal
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ blockSums, uint32_t G, uint32_t* __restrict__ dTotal) {
  __shared__ uint32_t ws[32];
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t v = t < G ? blockSums[t] : 0;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = ws[lane]; uint32_t y = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t z = __shfl_up_sync(0xffffffffu, y, o); if ((int)lane >= o) y += z; }
    ws[lane] = y - w;
    if (lane == 31) *dTotal = y;
  }
  __syncthreads();
  if (t < G) blockSums[t] = ws[warp] + x - v;
}
__global__ void __launch_bounds__(RS_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ dN, const uint32_t* __restrict__ blockSums) {
  __shared__ uint32_t ws[RS_WARPS];
  __shared__ uint32_t carry;
  uint32_t b, e; rs_chunk(*dN, b, e);
  if (threadIdx.x == 0) carry = blockSums[blockIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = b; base < e; base += RS_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < e ? in[i] : 0;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += t; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) { const uint32_t c = ws[w]; if (w < (int)warp) woff += c; tot += c; }
    if (i < e) out[i] = carry + woff + x - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
}
static inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, const uint32_t* dN, uint32_t* dTotal, uint32_t* blockSums, uint32_t ctas, cudaStream_t stream) {
  k_scan_reduce<<<ctas, RS_THREADS, 0, stream>>>(in, dN, blockSums);
  k_scan_sums<<<1, 1024, 0, stream>>>(blockSums, ctas, dTotal);  // ctas <= RS_MAX_CTAS <= 1024
  k_scan_apply<<<ctas, RS_THREADS, 0, stream>>>(in, out, dN, blockSums);
}
