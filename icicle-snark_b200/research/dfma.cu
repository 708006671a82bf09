// FP64-pipe probes for sm_100a and a prototype of a DFMA-based Montgomery product (research tool; built into
// lib/libicicle_b200_tools.so and, with -DDFMA_MAIN, as a standalone binary - never part of the product library).
//
// Why: the 32x32+64 multiply-add (IMAD.WIDE) issues once per 4 cycles per SM sub-partition on B200 (pipes2.cu), which
// bounds the 8x32-bit CIOS product at ~566 cycles per warp.  B200 (unlike B300) keeps a full-rate FP64 pipe; a 52x52-bit
// product split into exact high and low halves costs two round-toward-zero DFMAs and one DADD (the technique of Emmart,
// Zheng, Weems, "Faster modular exponentiation using double precision floating point arithmetic on the GPU", ARITH 2018),
// i.e. 2704 bit-products per ~6 FP64-pipe cycles against 1024 per 4 IMAD-pipe cycles.
//
//   D1 DFMA.RZ chains, D2 DADD chains, D3 DFMA + 64-bit integer add, D4 DFMA + IMAD.WIDE (co-issue of the two pipes),
//   M1 throughput of mul52 (5 x 52-bit limbs held as doubles, R = 2^260), checked on the host by dfma_check.py.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#define CH 8
template <int T>
__global__ void __launch_bounds__(256) dprobe(double* out, int iters, double s0)
{
  double x[CH], y = 1.0000001 + s0 * 1e-9, z = 0.999999 - s0 * 1e-9;
  uint64_t w[CH], v[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    x[k] = 1.0 + threadIdx.x * 1e-3 + k;
    w[k] = threadIdx.x * 2654435761u + k;
    v[k] = blockIdx.x * 40503u + k;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (T == 1) { // 2 DFMA.RZ
          x[k] = __fma_rz(x[k], y, z);
          x[k] = __fma_rz(x[k], z, y);
        }
        if (T == 2) { // 2 DADD
          x[k] = __dadd_rn(x[k], y);
          x[k] = __dadd_rn(x[k], z);
        }
        if (T == 3) { // 2 DFMA + 2 64-bit integer adds on other data
          x[k] = __fma_rz(x[k], y, z);
          x[k] = __fma_rz(x[k], z, y);
          asm volatile("add.u64 %0, %0, %1;" : "+l"(w[k]) : "l"(v[k]));
          asm volatile("add.u64 %0, %0, %1;" : "+l"(v[k]) : "l"(w[k]));
        }
        if (T == 4) { // 2 DFMA + 1 IMAD.WIDE accumulate
          x[k] = __fma_rz(x[k], y, z);
          x[k] = __fma_rz(x[k], z, y);
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, h, %1, %0;}" : "+l"(w[k]) : "r"((uint32_t)v[k]));
        }
        if (T == 5) { // the product pattern: DFMA, DADD, DFMA + two 64-bit adds of the bit patterns
          double hi = __fma_rz(x[k], y, 20282409603651670423947251286016.0);
          double d = __dsub_rn(20282409603651674927546878656512.0, hi);
          double lo = __fma_rz(x[k], y, d);
          w[k] += (uint64_t)__double_as_longlong(hi);
          v[k] += (uint64_t)__double_as_longlong(lo);
          x[k] = lo;
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k)
    s += x[k] + (double)(w[k] ^ v[k]);
  if (s == 1.2345) out[0] = s;
}

template <int T>
static double run_dcycles(double instr_per_u, float* ms_out = nullptr)
{
  cudaDeviceProp p;
  int dev = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&p, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int sms = p.multiProcessorCount, iters = 1000, blocks = sms * 8;
  double* d = nullptr;
  if (cudaMalloc(&d, 1024) != cudaSuccess) return -1;
  dprobe<T><<<blocks, 256>>>(d, 10, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    dprobe<T><<<blocks, 256>>>(d, iters, r + 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) return -1;
  if (ms_out) *ms_out = best;
  const double warps_per_smsp = (double)blocks * 8 / (sms * 4.0);
  const double instr = warps_per_smsp * iters * 4.0 * instr_per_u;
  return best * 1e-3 * khz * 1e3 / instr;
}

// ------------------------------------------------------------------------------------------------ mul52 prototype
// BN254 base field q in five 52-bit limbs; values are integer-valued doubles in [0, 2^52), little-endian limbs;
// Montgomery radix 2^260.  Inputs < 8q give an output < 2q (64 q^2 / 2^260 + q < 2q), so additions can stay lazy.
namespace f52 {
  __device__ __constant__ double kP[5];   // q limbs
  __device__ __constant__ uint64_t kNP;   // -q^-1 mod 2^52
  static const double C1 = 20282409603651670423947251286016.0;  // 2^104
  static const double C2 = 20282409603651674927546878656512.0;  // 2^104 + 2^52
  static const double K52 = 4503599627370496.0;                 // 2^52

  struct F {
    double l[5];
  };

  __device__ __forceinline__ void prod(double a, double b, uint64_t& lo_col, uint64_t& hi_col)
  {
    const double hi = __fma_rz(a, b, C1);
    const double d = __dsub_rn(C2, hi);
    const double lo = __fma_rz(a, b, d);
    hi_col += (uint64_t)__double_as_longlong(hi);
    lo_col += (uint64_t)__double_as_longlong(lo);
  }

  __device__ __forceinline__ F mul(const F& a, const F& b)
  {
    // column accumulators start at minus the exponent patterns they will collect
    const uint64_t EH = 0x467ull << 52, EL = 0x433ull << 52;
    uint64_t c[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      // products a_i b_j with i + j = k contribute lo here (k <= 8), with i + j = k - 1 their hi; the reduction adds
      // q_i p_j the same way for i + j = k
      const int nlo_ab = k <= 8 ? (k < 5 ? k + 1 : 9 - k) : 0;
      const int nhi_ab = k >= 1 ? (k - 1 < 5 ? k : 10 - k) : 0;
      c[k] = 0ull - (uint64_t)(2 * nlo_ab) * EL - (uint64_t)(2 * nhi_ab) * EH;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j)
        prod(a.l[i], b.l[j], c[i + j], c[i + j + 1]);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint64_t q = (c[i] * kNP) & 0xFFFFFFFFFFFFFull;
      const double qd = __dsub_rn(__longlong_as_double((long long)(q | EL)), K52);
#pragma unroll
      for (int j = 0; j < 5; ++j)
        prod(qd, kP[j], c[i + j], c[i + j + 1]);
      c[i + 1] += c[i] >> 52;
    }
    F r;
#pragma unroll
    for (int k = 5; k < 9; ++k) {
      c[k + 1] += c[k] >> 52;
      c[k] &= 0xFFFFFFFFFFFFFull;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
      r.l[k] = __dsub_rn(__longlong_as_double((long long)(c[5 + k] | EL)), K52);
    return r;
  }
} // namespace f52

template <int NCH>
__global__ void __launch_bounds__(128) mul52_kernel(const uint64_t* in, uint64_t* out, int iters)
{
  // in: per thread NCH x 2 x 5 limbs (integers); out: NCH x 5 limbs after `iters` steps of x = x*y
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  f52::F x[NCH], y[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      x[k].l[i] = (double)in[(t * NCH + k) * 10 + i];
      y[k].l[i] = (double)in[(t * NCH + k) * 10 + 5 + i];
    }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int k = 0; k < NCH; ++k)
      x[k] = f52::mul(x[k], y[k]);
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int i = 0; i < 5; ++i)
      out[(t * NCH + k) * 5 + i] = (uint64_t)x[k].l[i];
}

static void q_limbs(uint64_t* l)
{
  // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47 cut into 52-bit limbs
  const uint64_t w[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  for (int i = 0; i < 5; ++i) {
    const int bit = 52 * i, wd = bit / 64, sh = bit % 64;
    uint64_t v = w[wd] >> sh;
    if (sh && wd + 1 < 4) v |= w[wd + 1] << (64 - sh);
    l[i] = v & 0xFFFFFFFFFFFFFull;
  }
}

static int mul52_setup()
{
  uint64_t l[5];
  q_limbs(l);
  double d[5];
  for (int i = 0; i < 5; ++i)
    d[i] = (double)l[i];
  // -q^-1 mod 2^52 by Newton iteration on the low limb
  uint64_t inv = 1;
  for (int i = 0; i < 6; ++i)
    inv *= 2 - l[0] * inv;
  const uint64_t np = (0 - inv) & 0xFFFFFFFFFFFFFull;
  if (cudaMemcpyToSymbol(f52::kP, d, sizeof d) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(f52::kNP, &np, sizeof np) != cudaSuccess) return -1;
  return 0;
}

extern "C" {
double b200_dprobe_cycles(int id)
{
  switch (id) {
  case 1: return run_dcycles<1>(16);
  case 2: return run_dcycles<2>(16);
  case 3: return run_dcycles<3>(32);
  case 4: return run_dcycles<4>(24);
  case 5: return run_dcycles<5>(8); // per product
  }
  return -1;
}

// runs x = x*y `iters` times on `threads` threads x NCH chains from host inputs (10 limbs per chain), returns ms of the
// timed launch (after a warm-up launch) and the outputs (5 limbs per chain)
float b200_mul52_run(const uint64_t* in, uint64_t* out, int threads, int nch, int iters)
{
  if (mul52_setup() != 0) return -1.f;
  if (nch != 1 && nch != 2 && nch != 4) return -1.f;
  uint64_t *din = nullptr, *dout = nullptr;
  const size_t nin = (size_t)threads * nch * 10, nout = (size_t)threads * nch * 5;
  if (cudaMalloc(&din, nin * 8) != cudaSuccess || cudaMalloc(&dout, nout * 8) != cudaSuccess) return -1.f;
  cudaMemcpy(din, in, nin * 8, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms = -1.f;
  for (int r = 0; r < 2; ++r) {
    cudaEventRecord(e0);
    if (nch == 1) mul52_kernel<1><<<threads / 128, 128>>>(din, dout, iters);
    if (nch == 2) mul52_kernel<2><<<threads / 128, 128>>>(din, dout, iters);
    if (nch == 4) mul52_kernel<4><<<threads / 128, 128>>>(din, dout, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaMemcpy(out, dout, nout * 8, cudaMemcpyDeviceToHost);
  cudaFree(din);
  cudaFree(dout);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (cudaGetLastError() != cudaSuccess) return -1.f;
  return ms;
}
}

#ifdef DFMA_MAIN
#include <vector>
int main(int argc, char** argv)
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, khz);
  const char* names[] = {"", "D1 DFMA.RZ", "D2 DADD", "D3 2 DFMA + 2 add.u64 (per counted instr, 64-bit add = 2)",
                         "D4 2 DFMA + IMAD.WIDE (per instr)", "D5 product pattern DFMA,DADD,DFMA + 2 add.u64 (per product)"};
  for (int id = 1; id <= 5; ++id)
    printf("%-70s %7.3f cycles per warp-instr per SMSP\n", names[id], b200_dprobe_cycles(id));
  // mul52: correctness dump + throughput
  uint64_t ql[5];
  q_limbs(ql);
  for (int nch = 1; nch <= 4; nch *= 2) {
    const int threads = p.multiProcessorCount * 128 * 16, iters = 200;
    std::vector<uint64_t> in((size_t)threads * nch * 10), out((size_t)threads * nch * 5);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < in.size(); ++i) {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      uint64_t v = s & 0xFFFFFFFFFFFFFull;
      if (i % 5 == 4) v = s % (ql[4]); // value < q
      in[i] = v;
    }
    float ms = b200_mul52_run(in.data(), out.data(), threads, nch, iters);
    const double prods = (double)threads * nch * iters;
    printf("mul52 chains/thread=%d: %.3f ms, %.1f G products/s, %.0f cycles per warp-product per SMSP\n", nch, ms, prods / ms / 1e6,
           ms * 1e-3 * khz * 1e3 / (prods / 32 / (p.multiProcessorCount * 4)));
    if (nch == 1 && argc > 1) {
      // one-step dump for the host check: iters = 1 on the first 4096 threads
      std::vector<uint64_t> o1(4096 * 5);
      b200_mul52_run(in.data(), o1.data(), 4096, 1, 1);
      FILE* f = fopen(argv[1], "wb");
      if (f) {
        fwrite(in.data(), 8, 4096 * 10, f);
        fwrite(o1.data(), 8, 4096 * 5, f);
        fclose(f);
      }
    }
  }
  return 0;
}
#endif
