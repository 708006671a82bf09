// Instruction-throughput probes for the integer pipes of sm_100a (research tool: built into lib/libicicle_b200_tools.so,
// NOT part of the product library).  8 independent chains per thread, 256 threads/CTA, 8 CTAs/SM.
//   b200_probe_cycles(t) -> cycles per counted warp-instruction per SM sub-partition at the nominal SM clock
// The SASS mix of every probe is checked with research/sassmix.py (ptxas hoists loop-invariant products and rewrites
// mad.wide chains: round 1's "IMAD.WIDE peak" microbenchmark contained no IMAD.WIDE at all after optimisation, which
// is how an 18 T/s figure - really the 64-bit-add rate - ended up as the roofline denominator).  Measured on B200 at
// 1965 MHz (profiles/r02_pipe_probes.md): IMAD.WIDE.U32 with or without carry in/out 4.03-4.11 cycles (8 lanes/clk/SMSP
// = 9.2 T wide MAC/s per GPU), IMAD lo 2.03, IMAD.HI 4.06, IADD3 1.05 (split over two pipes), LOP3/SHF 2.03.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define MACW(acc, mul) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, h, %1, %0;}" : "+l"(acc) : "r"(mul))
template <int T>
__global__ void __launch_bounds__(256) probe(uint32_t* out, int iters, uint32_t s0)
{
  uint32_t x = threadIdx.x * 2654435761u + s0, y = (blockIdx.x + 7u) * 40503u + s0;
  uint32_t lo[CHAINS], hi[CHAINS];
  uint64_t w[CHAINS], v[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) {
    lo[k] = x + k;
    hi[k] = y + k;
    w[k] = ((uint64_t)hi[k] << 32) | lo[k];
    v[k] = ((uint64_t)lo[k] << 32) | hi[k];
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < CHAINS; ++k) {
        if (T == 1) { // 2 IMAD.WIDE accumulate form (multiplicands = halves of the accumulator)
          uint64_t p, q;
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, l, %2;}" : "=l"(p) : "l"(w[k]), "r"(y));
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, h, %2;}" : "=l"(q) : "l"(w[k]), "r"(x));
          asm volatile("{.reg .u64 t; add.u64 t, %0, %1; add.u64 %0, t, %2;}" : "+l"(w[k]) : "l"(p), "l"(q));
        }
        if (T == 2) { // same with immediate multipliers
          uint64_t p, q;
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, l, 0x1c72a34f;}" : "=l"(p) : "l"(w[k]));
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, h, 0x2d522d07;}" : "=l"(q) : "l"(w[k]));
          asm volatile("{.reg .u64 t; add.u64 t, %0, %1; add.u64 %0, t, %2;}" : "+l"(w[k]) : "l"(p), "l"(q));
        }
        if (T == 4) { // plain IADD3, no carry
          asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[k]) : "r"(hi[k]));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[k]) : "r"(lo[k]));
        }
        if (T == 5) // 64-bit add = IADD3 (carry out) + IADD3.X
          asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(hi[(k + 1) % CHAINS]), "r"(lo[(k + 3) % CHAINS]));
        if (T == 6) { // 3-way 64-bit add: IADD3 with two carry-outs + IADD3.X with two carry-ins
          asm volatile("{.reg .u64 t; add.u64 t, %0, %1; add.u64 %0, t, %2;}" : "+l"(w[k]) : "l"(w[(k + 1) % CHAINS]), "l"(v[k]));
        }
        if (T == 9) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[k]) : "r"(hi[k]), "r"(y));
        if (T == 10) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[k]) : "r"(hi[k]), "r"(y));
        if (T == 11) { // 2 IMAD.WIDE accumulate + 2 plain IADD3 on other data
          uint64_t p, q;
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, l, %2;}" : "=l"(p) : "l"(w[k]), "r"(y));
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, h, %2;}" : "=l"(q) : "l"(w[k]), "r"(x));
          asm volatile("{.reg .u64 t; add.u64 t, %0, %1; add.u64 %0, t, %2;}" : "+l"(w[k]) : "l"(p), "l"(q));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[k]) : "r"(hi[k]));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[k]) : "r"(lo[k]));
        }
        if (T == 12) { // 2 IMAD.WIDE accumulate + one 64-bit carry add (2 ALU) on other data
          uint64_t p, q;
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, l, %2;}" : "=l"(p) : "l"(w[k]), "r"(y));
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, h, %2;}" : "=l"(q) : "l"(w[k]), "r"(x));
          asm volatile("{.reg .u64 t; add.u64 t, %0, %1; add.u64 %0, t, %2;}" : "+l"(w[k]) : "l"(p), "l"(q));
          asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(hi[(k + 1) % CHAINS]), "r"(lo[(k + 3) % CHAINS]));
        }
        if (T == 13) asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(lo[k]) : "r"(hi[k]));
        if (T == 14) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[k]) : "r"(hi[k]), "r"(y));
        if (T == 17) { // 2 IMAD.WIDE accumulate + one 3-way 64-bit add on other data
          uint64_t p, q;
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, l, %2;}" : "=l"(p) : "l"(w[k]), "r"(y));
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %1; mul.wide.u32 %0, h, %2;}" : "=l"(q) : "l"(w[k]), "r"(x));
          asm volatile("{.reg .u64 t; add.u64 t, %0, %1; add.u64 %0, t, %2;}" : "+l"(w[k]) : "l"(p), "l"(q));
          uint64_t vv = ((uint64_t)hi[k] << 32) | lo[k], v2 = ((uint64_t)lo[(k + 1) % CHAINS] << 32) | hi[(k + 2) % CHAINS];
          asm volatile("{.reg .u64 t; add.u64 t, %0, %1; add.u64 %0, t, %2;}" : "+l"(v[k]) : "l"(vv), "l"(v2));
        }
      }
      if (T == 7) { // 8-limb add with carry chain (256-bit add of two varying arrays), twice
        asm volatile(
          "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
          : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7])
          : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]));
        asm volatile(
          "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
          : "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7])
          : "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]));
      }
      if (T == 8 || T == 18) { // IMAD.WIDE.X carry chain: 8 wide MACs (18: plus two 8-limb carry-add chains)
        asm volatile(
          "mad.lo.cc.u32 %0, %16, %1, %0;\n\tmadc.hi.cc.u32 %1, %16, %1, %1;\n\t"
          "madc.lo.cc.u32 %2, %16, %3, %2;\n\tmadc.hi.cc.u32 %3, %16, %3, %3;\n\t"
          "madc.lo.cc.u32 %4, %16, %5, %4;\n\tmadc.hi.cc.u32 %5, %16, %5, %5;\n\t"
          "madc.lo.cc.u32 %6, %16, %7, %6;\n\tmadc.hi.cc.u32 %7, %16, %7, %7;\n\t"
          "madc.lo.cc.u32 %8, %16, %9, %8;\n\tmadc.hi.cc.u32 %9, %16, %9, %9;\n\t"
          "madc.lo.cc.u32 %10, %16, %11, %10;\n\tmadc.hi.cc.u32 %11, %16, %11, %11;\n\t"
          "madc.lo.cc.u32 %12, %16, %13, %12;\n\tmadc.hi.cc.u32 %13, %16, %13, %13;\n\t"
          "madc.lo.cc.u32 %14, %16, %15, %14;\n\tmadc.hi.u32 %15, %16, %15, %15;\n\t"
          : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
            "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
          : "r"(y));
      }
      if (T == 18) {
        uint32_t* a = reinterpret_cast<uint32_t*>(w);
        uint32_t* b = reinterpret_cast<uint32_t*>(v);
        asm volatile(
          "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
          : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
          : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
        asm volatile(
          "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
          : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k)
    s ^= lo[k] ^ hi[k] ^ (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32) ^ (uint32_t)v[k] ^ (uint32_t)(v[k] >> 32);
  if (s == 0x12345678u) out[0] = s;
}

template <int T>
static double run_cycles(double instr_per_u)
{
  cudaDeviceProp p;
  int dev = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&p, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int sms = p.multiProcessorCount, iters = 1000, blocks = sms * 8;
  uint32_t* d = nullptr;
  if (cudaMalloc(&d, 1024) != cudaSuccess) return -1;
  probe<T><<<blocks, 256>>>(d, 10, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<T><<<blocks, 256>>>(d, iters, r + 2);
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
  const double warps_per_smsp = (double)blocks * 8 / (sms * 4.0);
  const double instr = warps_per_smsp * iters * 4.0 * instr_per_u;
  return best * 1e-3 * khz * 1e3 / instr;
}

struct ProbeDesc {
  int id;
  const char* name;
  double per_u;
};
static const ProbeDesc kProbes[] = {
  {1, "IMAD.WIDE.U32 R,R,R,R (64-bit accumulate, no carry)", 16},
  {2, "IMAD.WIDE.U32 R,R,imm,R", 16},
  {4, "IADD3 / IMAD.IADD, no carry", 16},
  {5, "64-bit add: IADD3 carry-out + IADD3.X", 16},
  {6, "3-way 64-bit add: dual-carry IADD3 + IADD3.X", 16},
  {7, "256-bit add, carry chain (per IADD3.X)", 16},
  {8, "IMAD.WIDE.U32.X carry chain (per wide MAC)", 8},
  {9, "IMAD lo (32-bit)", 8},
  {10, "IMAD.HI.U32", 8},
  {11, "2 IMAD.WIDE + 2 IADD3 (per instruction)", 32},
  {12, "2 IMAD.WIDE + 64-bit carry add (per instruction)", 32},
  {17, "2 IMAD.WIDE + 3-way 64-bit add (per instruction)", 32},
  {18, "8 IMAD.WIDE.X + 16 carry-chain IADD3.X (per instruction)", 24},
  {13, "SHF funnel shift", 8},
  {14, "LOP3", 8},
};

extern "C" {
// cycles per counted warp-instruction per SMSP for probe `id` (see kProbes), < 0 on error / unknown id
double b200_probe_cycles(int id)
{
  switch (id) {
  case 1: return run_cycles<1>(16);
  case 2: return run_cycles<2>(16);
  case 4: return run_cycles<4>(16);
  case 5: return run_cycles<5>(16);
  case 6: return run_cycles<6>(16);
  case 7: return run_cycles<7>(16);
  case 8: return run_cycles<8>(8);
  case 9: return run_cycles<9>(8);
  case 10: return run_cycles<10>(8);
  case 11: return run_cycles<11>(32);
  case 12: return run_cycles<12>(32);
  case 17: return run_cycles<17>(32);
  case 18: return run_cycles<18>(24);
  case 13: return run_cycles<13>(8);
  case 14: return run_cycles<14>(8);
  }
  return -1;
}
int b200_probe_count(void) { return (int)(sizeof(kProbes) / sizeof(kProbes[0])); }
int b200_probe_id(int k) { return kProbes[k].id; }
const char* b200_probe_name(int k) { return kProbes[k].name; }
// wide (32x32+64) multiply-adds per second the whole GPU sustains: the integer-pipe roofline denominator of the MSM / NTT
double b200_imad_wide_peak(void)
{
  cudaDeviceProp p;
  int dev = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&p, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double cyc = run_cycles<1>(16);
  if (!(cyc > 0)) return -1;
  return (double)p.multiProcessorCount * 4 * 32 * (khz * 1e3) / cyc;
}
}

#ifdef PIPES2_MAIN
int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%s sms=%d clock=%d kHz (cycles computed at this nominal clock)\n", p.name, p.multiProcessorCount, khz);
  for (int k = 0; k < b200_probe_count(); ++k)
    printf("T%-3d %-58s %6.3f cyc per counted warp-instr per SMSP\n", kProbes[k].id, kProbes[k].name, b200_probe_cycles(kProbes[k].id));
  printf("IMAD.WIDE peak: %.2f T MAC/s\n", b200_imad_wide_peak() / 1e12);
  return 0;
}
#endif
