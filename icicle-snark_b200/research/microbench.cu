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
    if (MODE == 6 || MODE == 12 || MODE == 13) {
      Fq x[4], y;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          x[k].v[i] = (threadIdx.x * 2654435761u + seed + i * 40503u + k) & 0x0fffffffu;
      y = x[0];
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (MODE == 6) x[k] = x[k] * y;
          #ifdef __CUDA_ARCH__
          if (MODE == 12) x[k] = x[k].sqr_sos(); // dedicated squaring (36 + 72 wide MACs)
#endif
#ifdef __CUDA_ARCH__
          if (MODE == 13) {                  // wide product + separate SOS reduction (64 + 72 wide MACs)
            uint32_t T[16];
            Fq::mul_wide(T, x[k].v, y.v);
            x[k] = Fq::redc_wide(T);
          }
#endif
        }
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


  // ---- FP64-pipe candidate: 6 x 48-bit limbs held as exact-integer doubles, Montgomery R = 2^288.
  // Every partial product x*y (< 2^96) is added to its column with a fixed-exponent chain:
  //   s' = fma_rz(x, y, s)            s in [2^100, 2^101): keeps the multiples of 2^48 (ulp), drops r = (s + xy) - s'
  //   r  = fma(x, y, s - s')          exact, in [0, 2^48)
  //   l += r                          exact while l < 2^53 (12 terms + carry)
  // so a column is (s - 2^100) + l with NO integer instructions: the multiplier lives on the FP64 pipe, which the
  // integer kernels leave idle and which co-issues with IMAD at full rate (mode 4).
  struct D48 {
    double v[6];
  };
  __device__ __forceinline__ void d48_addprod(double& s, double& l, double x, double y)
  {
    double sn = __fma_rz(x, y, s);
    double d = __dsub_rn(s, sn);
    double r = __fma_rn(x, y, d);
    l = __dadd_rn(l, r);
    s = sn;
  }
  __device__ __forceinline__ D48 mul48(const D48& a, const D48& b)
  {
    const double C100 = 1267650600228229401496703205376.0;  // 2^100
    const double I48 = 3.552713678800501e-15;               // 2^-48
    const double NINV = (double)0x782e4866389ull;           // -q^-1 mod 2^48
    const double PL[6] = {(double)0x8c16d87cfd47ull, (double)0x6871ca8d3c20ull, (double)0x585d97816a91ull,
                          (double)0xb85045b68181ull, (double)0x4e72e131a029ull, (double)0x3064ull};
    double s[12], l[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      s[k] = C100;
      l[k] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j < 6; ++j)
        d48_addprod(s[i + j], l[i + j], a.v[j], b.v[i]);
      double u = __dadd_rz(l[i], C100);
      double tl = __dsub_rn(l[i], __dsub_rn(u, C100)); // low 48 bits of the column
      double hq = __fma_rz(tl, NINV, C100);
      double q = __fma_rn(tl, NINV, __dsub_rn(C100, hq)); // tl * n' mod 2^48
#pragma unroll
      for (int j = 0; j < 6; ++j)
        d48_addprod(s[i + j], l[i + j], q, PL[j]);
      double carry = __fma_rn(l[i], I48, __dmul_rn(__dsub_rn(s[i], C100), I48));
      l[i + 1] = __dadd_rn(l[i + 1], carry);
    }
    D48 r;
#pragma unroll
    for (int k = 6; k < 12; ++k) {
      double u = __dadd_rz(l[k], C100);
      double hi = __dsub_rn(u, C100);
      r.v[k - 6] = __dsub_rn(l[k], hi);
      if (k < 11) {
        double carry = __fma_rn(hi, I48, __dmul_rn(__dsub_rn(s[k], C100), I48));
        l[k + 1] = __dadd_rn(l[k + 1], carry);
      }
    }
    return r;
  }
  __device__ __forceinline__ D48 to48(const Fq& x)
  {
    D48 r;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      int bit = 48 * k, w = bit >> 5, sh = bit & 31; // sh is 0 or 16
      uint64_t lo = x.v[w], mid = w + 1 < 8 ? x.v[w + 1] : 0, top = w + 2 < 8 ? x.v[w + 2] : 0;
      uint64_t val = sh == 0 ? (lo | ((mid & 0xffffull) << 32)) : ((lo >> 16) | (mid << 16));
      (void)top;
      val &= 0xffffffffffffull;
      r.v[k] = __longlong_as_double((long long)(val | 0x4330000000000000ull)) - 4503599627370496.0;
    }
    return r;
  }
  __device__ __forceinline__ Fq from48(const D48& x) // limbs normalised (< 2^48), value < 2^256
  {
    uint64_t l[6];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      l[k] = (uint64_t)__double_as_longlong(x.v[k] + 4503599627370496.0) & 0xffffffffffffull;
    Fq r;
    r.v[0] = (uint32_t)l[0];
    r.v[1] = (uint32_t)(l[0] >> 32) | (uint32_t)(l[1] << 16);
    r.v[2] = (uint32_t)(l[1] >> 16);
    r.v[3] = (uint32_t)l[2];
    r.v[4] = (uint32_t)(l[2] >> 32) | (uint32_t)(l[3] << 16);
    r.v[5] = (uint32_t)(l[3] >> 16);
    r.v[6] = (uint32_t)l[4];
    r.v[7] = (uint32_t)(l[4] >> 32) | (uint32_t)(l[5] << 16);
    return r;
  }

  // self-check: 2^32 * mul48(a,b) == CIOS(a,b) mod q  (R = 2^288 vs 2^256)
  __global__ void mul48_check_kernel(uint32_t seed, int* bad)
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
    a.v[7] &= 0x1fffffffu;
    b.v[7] &= 0x1fffffffu;
    if (threadIdx.x == 0) { // edge: q - 1
      for (int i = 0; i < 8; ++i)
        a.v[i] = FqCfg::P(i);
      a.v[0] -= 1;
    }
    Fq want = a * b;
    Fq got = from48(mul48(to48(a), to48(b)));
    Fq zero = Fq::zero();
    got = got + zero; // < 2q -> canonical
    for (int k = 0; k < 32; ++k)
      got = got.dbl();
    // and a dependent chain through the FP64 representation
    D48 x = to48(a), y = to48(b);
    Fq w2 = a;
    for (int k = 0; k < 5; ++k) {
      x = mul48(x, y);
      w2 = w2 * b;
    }
    Fq g2 = from48(x) + zero;
    for (int k = 0; k < 160; ++k)
      g2 = g2.dbl();
    if (got != want || g2 != w2) atomicAdd(bad, 1);
  }

  // 10: FP64 multiplier alone; 11: co-issue test - 256-thread CTAs, warps 0-3 run the integer CIOS (3/2 x the
  // iterations, matching its speed) and warps 4-7 the FP64 multiplier, so every SM sub-partition (warp id mod 4)
  // holds both kinds
  template <int MODE>
  __global__ void __launch_bounds__(256) fieldmul48_kernel(uint32_t* out, int iters, uint32_t seed)
  {
    const bool use_fp = MODE == 10 || ((threadIdx.x >> 7) & 1);
    if (MODE == 11 && !use_fp) iters = iters * 3 / 2;
    Fq x[4], y;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x[k].v[i] = (threadIdx.x * 2654435761u + seed + i * 40503u + k) & 0x0fffffffu;
    y = x[0];
    uint32_t sacc = 0;
    if (use_fp) {
      D48 dx[4], dy = to48(y);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        dx[k] = to48(x[k]);
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          dx[k] = mul48(dx[k], dy);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 6; ++i)
          sacc ^= (uint32_t)__double_as_longlong(dx[k].v[i]);
    } else {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          x[k] = x[k] * y;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sacc ^= x[k].v[i];
    }
    if (sacc == 0x12345678u) out[0] = sacc;
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
    if (mode == 10 || mode == 11) {
      const int it2 = 256, bl2 = sm_count() * 4;
      if (mode == 10)
        fieldmul48_kernel<10><<<bl2, 256>>>((uint32_t*)d, it2, rep);
      else
        fieldmul48_kernel<11><<<bl2, 256>>>((uint32_t*)d, it2, rep);
      cudaEventRecord(e1, 0);
      if (cudaEventSynchronize(e1) != cudaSuccess) break;
      float msx = 0;
      cudaEventElapsedTime(&msx, e0, e1);
      // mode 11: half the threads do it2*3/2 iterations, half it2
      double per_thread = mode == 10 ? (double)it2 * 4 : ((double)(it2 * 3 / 2) + it2) * 4 / 2;
      double rx = (double)bl2 * 256 * per_thread / (msx * 1e-3);
      if (rx > best) best = rx;
      continue;
    }
    if (mode >= 6) {
      const int it2 = 512, bl2 = sm_count() * 8;
      if (mode == 6)
        fieldmul_kernel<6><<<bl2, 128>>>((uint32_t*)d, it2, rep);
      else if (mode == 7)
        fieldmul_kernel<7><<<bl2, 128>>>((uint32_t*)d, it2, rep);
      else if (mode == 12)
        fieldmul_kernel<12><<<bl2, 128>>>((uint32_t*)d, it2, rep);
      else if (mode == 13)
        fieldmul_kernel<13><<<bl2, 128>>>((uint32_t*)d, it2, rep);
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

// Co-residency test with separately compiled kernels (each keeps its own register budget), two streams:
// returns rates (field products/s): out[0] integer CIOS alone, out[1] FP64 multiplier alone, out[2] both together.
extern "C" int b200_corun_test(double* out3, int ctas_int_per_sm, int ctas_fp_per_sm)
{
  if (ensure_device() != ICICLE_SUCCESS) return -1;
  uint32_t* d = nullptr;
  if (cudaMalloc((void**)&d, 64) != cudaSuccess) return -1;
  cudaStream_t sa, sb;
  cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
  cudaEvent_t e0, ea, eb;
  cudaEventCreate(&e0);
  cudaEventCreate(&ea);
  cudaEventCreate(&eb);
  const int ga = sm_count() * ctas_int_per_sm, gb = sm_count() * ctas_fp_per_sm;
  const int ita = 1536, itb = 1024; // ~ equal duration alone at 4:2 CTAs
  for (int mode = 0; mode < 3; ++mode) {
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, sa);
      cudaStreamWaitEvent(sb, e0, 0);
      if (mode != 1) fieldmul_kernel<6><<<ga, 128, 0, sa>>>(d, ita, rep);
      if (mode != 0) fieldmul48_kernel<10><<<gb, 128, 0, sb>>>(d, itb, rep);
      cudaEventRecord(ea, sa);
      cudaEventRecord(eb, sb);
      cudaEventSynchronize(ea);
      cudaEventSynchronize(eb);
      float ta = 0, tb = 0;
      cudaEventElapsedTime(&ta, e0, ea);
      cudaEventElapsedTime(&tb, e0, eb);
      float t = ta > tb ? ta : tb;
      double muls = (mode != 1 ? (double)ga * 128 * ita * 4 : 0) + (mode != 0 ? (double)gb * 128 * itb * 4 : 0);
      double r = muls / (t * 1e-3);
      if (r > best) best = r;
    }
    out3[mode] = best;
  }
  cudaFree(d);
  cudaStreamDestroy(sa);
  cudaStreamDestroy(sb);
  return 0;
}

extern "C" int b200_mul48_selfcheck(void)
{
  if (ensure_device() != ICICLE_SUCCESS) return -1;
  int* d = nullptr;
  if (cudaMalloc((void**)&d, 4) != cudaSuccess) return -1;
  cudaMemset(d, 0, 4);
  mul48_check_kernel<<<64, 128>>>(777u, d);
  int h = -1;
  cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return h;
}

// the roofline denominator used by bench.py: wide (32x32+64) multiply-adds per second in carry chains
extern "C" double b200_imad_peak(int wide) { return b200_pipe_peak(wide ? 5 : 0); }
