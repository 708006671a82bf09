// Arithmetic-pipe microbenchmarks: the roofline denominators for MSM/NTT ("measure with an
// IMAD microbenchmark like MEASURED_PEAKS does for HBM"). Independent register chains, no memory traffic.
//   mode 0: mad.lo.u32   (IMAD)            mode 1: mad.wide.u32 (IMAD.WIDE, 32x32+64)
//   mode 2: mad.hi.u32   (IMAD.HI)         mode 3: fma.rn.f64   (DFMA)
//   mode 4: 4 IMAD.WIDE chains + 4 DFMA chains interleaved (do the two pipes issue concurrently?)
//   mode 5: mad.lo.cc/madc.hi pairs as the field multiplier emits them (IMAD.WIDE.U32.X carry chains)
#include "common.cuh"
#include "field.cuh"

namespace b200 {

  template <int MODE>
  __global__ void __launch_bounds__(256) pipe_kernel(uint64_t* out, int iters, uint32_t seed)
  {
    uint32_t b = (threadIdx.x * 2654435761u + seed) | 1u, c = blockIdx.x + 12345u;
    uint64_t sink = 0;
    if (MODE == 0 || MODE == 2) {
      uint32_t a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        a[k] = threadIdx.x + k;
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (MODE == 0)
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
          else
            asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sink ^= a[k];
    } else if (MODE == 1) {
      uint64_t a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        a[k] = threadIdx.x + k;
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(b), "r"(c));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sink ^= a[k];
    } else if (MODE == 3) {
      double a[8], fb = 1.0 + 1e-9 * b, fc = 1e-3 * c;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        a[k] = threadIdx.x + k;
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(fb), "d"(fc));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sink ^= (uint64_t)__double_as_longlong(a[k]);
    } else if (MODE == 4) {
      uint64_t a[4];
      double d[4], fb = 1.0 + 1e-9 * b, fc = 1e-3 * c;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a[k] = threadIdx.x + k;
        d[k] = threadIdx.x + k;
      }
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(b), "r"(c));
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(fb), "d"(fc));
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sink ^= a[k] ^ (uint64_t)__double_as_longlong(d[k]);
    } else {
      uint32_t lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lo[k] = threadIdx.x + k;
        hi[k] = k;
      }
      for (int i = 0; i < iters; ++i) {
        // one 8-product carry chain per iteration, like a multiplier row: counts as 8 wide multiply-adds
        asm volatile(
          "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
          "madc.lo.cc.u32 %2, %16, %18, %2;\n\tmadc.hi.cc.u32 %3, %16, %18, %3;\n\t"
          "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
          "madc.lo.cc.u32 %6, %16, %18, %6;\n\tmadc.hi.cc.u32 %7, %16, %18, %7;\n\t"
          "madc.lo.cc.u32 %8, %16, %17, %8;\n\tmadc.hi.cc.u32 %9, %16, %17, %9;\n\t"
          "madc.lo.cc.u32 %10, %16, %18, %10;\n\tmadc.hi.cc.u32 %11, %16, %18, %11;\n\t"
          "madc.lo.cc.u32 %12, %16, %17, %12;\n\tmadc.hi.cc.u32 %13, %16, %17, %13;\n\t"
          "madc.lo.cc.u32 %14, %16, %18, %14;\n\tmadc.hi.u32 %15, %16, %18, %15;\n\t"
          : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
            "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
          : "r"(b), "r"(c), "r"(seed));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sink ^= lo[k] ^ ((uint64_t)hi[k] << 32);
    }
    if (sink == 0x123456789abcdefull) out[0] = sink;
  }

  // ---- field-multiplier candidates (throughput of dependent Montgomery products, 4 chains per thread)
  // mode 6: the production 8x32-bit CIOS (carry chains, IMAD.WIDE.X); mode 7: 9x29-bit limbs with plain
  // IMAD.WIDE column accumulation (no carry-in), lazy carries.
  struct F29 {
    uint32_t v[9];
  };
  // 9x29-bit Montgomery product, R = 2^261, operand scanning with a sliding window of nine 64-bit columns.
  // Every partial product is a plain IMAD.WIDE (64-bit accumulate, no carry in/out): limbs < 2^30 keep a column
  // below 2^64 over its 18 terms.  Output limbs < 2^29 (top limb slightly more), value < 2p for inputs < 4p.
  template <bool CONST_IN_REGS>
  __device__ __forceinline__ F29 mul29(const F29& a, const F29& b)
  {
    constexpr uint32_t MASK = (1u << 29) - 1, PINV = 0x4866389u;
    constexpr uint32_t PLc[9] = {0x187cfd47, 0x10460b6, 0x1c72a34f, 0x2d522d0, 0x1585d978, 0x2db40c0, 0xa6e141, 0xe5c2634, 0x30644e};
    uint32_t PL[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      PL[j] = PLc[j];
      if (CONST_IN_REGS) asm volatile("mov.u32 %0, %1;" : "=r"(PL[j]) : "r"(PLc[j])); // keep the modulus in registers
    }
    uint64_t c[10];
#pragma unroll
    for (int j = 0; j < 9; ++j)
      asm("mul.wide.u32 %0, %1, %2;" : "=l"(c[j]) : "r"(a.v[j]), "r"(b.v[0]));
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (i > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[j]) : "r"(a.v[j]), "r"(b.v[i]));
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(c[8]) : "r"(a.v[8]), "r"(b.v[i]));
      }
      uint32_t m = ((uint32_t)c[0] * PINV) & MASK;
#pragma unroll
      for (int j = 0; j < 9; ++j)
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[j]) : "r"(m), "r"(PL[j]));
      uint64_t carry = c[0] >> 29;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        c[j] = c[j + 1];
      c[0] += carry;
    }
    F29 r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      uint64_t v = c[k] + carry;
      r.v[k] = k < 8 ? (uint32_t)v & MASK : (uint32_t)v;
      carry = v >> 29;
    }
    return r;
  }

  // 8x32 (value < 2^256) <-> 9x29
  __device__ __forceinline__ F29 to29(const Fq& x)
  {
    F29 r;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      int bit = 29 * k, w = bit >> 5, sh = bit & 31;
      uint64_t two = x.v[w];
      if (w + 1 < 8) two |= (uint64_t)x.v[w + 1] << 32;
      r.v[k] = (uint32_t)(two >> sh) & ((1u << 29) - 1);
    }
    return r;
  }
  __device__ __forceinline__ Fq from29(const F29& x) // value must be < 2^256
  {
    Fq r = Fq::zero();
    uint64_t acc = 0;
    int filled = 0, w = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      acc |= (uint64_t)x.v[k] << filled;
      filled += 29;
      if (filled >= 32 && w < 8) {
        r.v[w++] = (uint32_t)acc;
        acc >>= 32;
        filled -= 32;
      }
    }
    if (w < 8) r.v[w] = (uint32_t)acc;
    return r;
  }

  // self-check: 32 * mul29(a,b) == CIOS(a,b) mod q  (R differs by 2^5). Returns mismatches in *bad.
  __global__ void mul29_check_kernel(uint32_t seed, int* bad)
  {
    uint32_t st = seed + blockIdx.x * 977u + threadIdx.x * 131u;
    Fq a, b;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      st = st * 1664525u + 1013904223u;
      a.v[i] = st;
      st = st * 1664525u + 1013904223u;
      b.v[i] = st;
    }
    a.v[7] &= 0x1fffffffu; // < 2^253 < q
    b.v[7] &= 0x1fffffffu;
    Fq want = a * b;
    F29 y = mul29<true>(to29(a), to29(b));
    F29 y2 = mul29<false>(to29(a), to29(b));
    Fq got = from29(y), got2 = from29(y2);
    // canonical reduce (result < 2q), then x32
    Fq zero = Fq::zero();
    got = got + zero;
    got2 = got2 + zero;
    for (int k = 0; k < 5; ++k) {
      got = got.dbl();
      got2 = got2.dbl();
    }
    if (got != want || got2 != want) atomicAdd(bad, 1);
  }

  template <int MODE>
  __global__ void __launch_bounds__(128) fieldmul_kernel(uint32_t* out, int iters, uint32_t seed)
  {
    if (MODE == 6) {
      Fq x[4], y;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          x[k].v[i] = (threadIdx.x * 2654435761u + seed + i * 40503u + k) & 0x0fffffffu;
      y = x[0];
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          x[k] = x[k] * y;
      }
      uint32_t s = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          s ^= x[k].v[i];
      if (s == 0x12345678u) out[0] = s;
    } else {
      F29 x[4], y;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 9; ++i)
          x[k].v[i] = (threadIdx.x * 2654435761u + seed + i * 40503u + k) & 0x0fffffffu;
      y = x[0];
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          x[k] = mul29<MODE == 8>(x[k], y);
      }
      uint32_t s = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 9; ++i)
          s ^= x[k].v[i];
      if (s == 0x12345678u) out[0] = s;
    }
  }

  // mode 9: "carry-save" wide multiply-add: IMAD.WIDE with carry-OUT only (no carry-in) + a counter bumped by the
  // carry on the ALU pipe - the building block of a multiplier without IMAD.WIDE.X chains
  __global__ void __launch_bounds__(256) carrysave_kernel(uint64_t* out, int iters, uint32_t seed)
  {
    uint32_t b = (threadIdx.x * 2654435761u + seed) | 1u, c = blockIdx.x + 12345u;
    uint64_t acc[8];
    uint32_t cnt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] = threadIdx.x + k;
      cnt[k] = 0;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        asm volatile(
          "{ .reg .u32 l, h;\n\t"
          "mov.b64 {l, h}, %0;\n\t"
          "mad.lo.cc.u32 l, %2, %3, l;\n\t"
          "madc.hi.cc.u32 h, %2, %3, h;\n\t"
          "addc.u32 %1, %1, 0;\n\t"
          "mov.b64 %0, {l, h}; }"
          : "+l"(acc[k]), "+r"(cnt[k])
          : "r"(b), "r"(c));
    }
    uint64_t sink = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      sink ^= acc[k] + cnt[k];
    if (sink == 0x123456789abcdefull) out[0] = sink;
  }

  template <int MODE>
  static void launch_pipe(uint64_t* d, int iters, int blocks, int rep)
  {
    pipe_kernel<MODE><<<blocks, 256>>>(d, iters, rep);
  }

} // namespace b200

using namespace b200;

// measured operations per second for the given mode on the active device (mode 4 counts both kinds); <0 on error
extern "C" double b200_pipe_peak(int mode)
{
  if (ensure_device() != ICICLE_SUCCESS) return -1.0;
  uint64_t* d = nullptr;
  if (cudaMalloc((void**)&d, 64) != cudaSuccess) return -1.0;
  const int iters = 4096, blocks = sm_count() * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = -1.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, 0);
    if (mode == 9) {
      carrysave_kernel<<<blocks, 256>>>(d, iters, rep);
      cudaEventRecord(e1, 0);
      if (cudaEventSynchronize(e1) != cudaSuccess) break;
      float ms9 = 0;
      cudaEventElapsedTime(&ms9, e0, e1);
      double r9 = (double)blocks * 256 * iters * 8 / (ms9 * 1e-3);
      if (r9 > best) best = r9;
      continue;
    }
    if (mode >= 6) {
      const int it2 = 512, bl2 = sm_count() * 8;
      if (mode == 6)
        fieldmul_kernel<6><<<bl2, 128>>>((uint32_t*)d, it2, rep);
      else if (mode == 7)
        fieldmul_kernel<7><<<bl2, 128>>>((uint32_t*)d, it2, rep);
      else
        fieldmul_kernel<8><<<bl2, 128>>>((uint32_t*)d, it2, rep);
      cudaEventRecord(e1, 0);
      if (cudaEventSynchronize(e1) != cudaSuccess) break;
      float ms2 = 0;
      cudaEventElapsedTime(&ms2, e0, e1);
      double rate2 = (double)bl2 * 128 * it2 * 4 / (ms2 * 1e-3); // field products per second
      if (rate2 > best) best = rate2;
      continue;
    }
    switch (mode) {
    case 0: launch_pipe<0>(d, iters, blocks, rep); break;
    case 1: launch_pipe<1>(d, iters, blocks, rep); break;
    case 2: launch_pipe<2>(d, iters, blocks, rep); break;
    case 3: launch_pipe<3>(d, iters, blocks, rep); break;
    case 4: launch_pipe<4>(d, iters, blocks, rep); break;
    default: launch_pipe<5>(d, iters, blocks, rep); break;
    }
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * iters * 8;
    double rate = ops / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

// number of mismatching products between the 9x29 candidate and the production multiplier (0 = agree)
extern "C" int b200_mul29_selfcheck(void)
{
  if (ensure_device() != ICICLE_SUCCESS) return -1;
  int* d = nullptr;
  if (cudaMalloc((void**)&d, 4) != cudaSuccess) return -1;
  cudaMemset(d, 0, 4);
  mul29_check_kernel<<<64, 128>>>(12345u, d);
  int h = -1;
  cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return h;
}

// the roofline denominator used by bench.py: wide (32x32+64) multiply-adds per second in carry chains
extern "C" double b200_imad_peak(int wide) { return b200_pipe_peak(wide ? 5 : 0); }
