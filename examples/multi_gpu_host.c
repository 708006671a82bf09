/* A host in plain C that shards one Groth16 proof over the GPUs of a box with the library's own data plane
 * (include/icicle_b200.h: b200_comm_*, b200_zkey_cache_create_sharded, b200_groth16_prove_sharded) - no Python, no MPI:
 * one process per GPU (fork), the 128-byte rendezvous token travels over pipes.  The reference has no multi-GPU path
 * (device 0 is hard-coded, src/lib.rs:29); this is what a Rust / Go / Java host would do through its FFI.
 *
 *   gcc -std=c99 -Iinclude examples/multi_gpu_host.c -o multi_gpu_host -Licicle-snark_b200/lib -licicle_b200 \
 *       -Wl,-rpath,$PWD/icicle-snark_b200/lib
 *   ./multi_gpu_host circuit_final.zkey witness.wtns proof.json <n_gpus>
 *
 * tests/test_abi.py compiles and links this file on the CPU box; it needs n_gpus B200s to run. */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include "icicle_b200.h"

static unsigned char* read_file(const char* path, size_t* len)
{
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  unsigned char* buf = (unsigned char*)malloc(n > 0 ? (size_t)n : 1);
  if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) {
    free(buf);
    buf = NULL;
  }
  fclose(f);
  *len = (size_t)n;
  return buf;
}

static int run_rank(int rank, int world, const unsigned char* token, const char* zkey_path, const char* wtns_path, const char* proof_path)
{
  icicleDevice dev;
  memset(&dev, 0, sizeof dev);
  strcpy(dev.type, "CUDA");
  dev.id = rank;
  if (icicle_set_device(&dev) != ICICLE_SUCCESS) return 10;
  b200_comm* comm = NULL;
  if (b200_comm_create(token, rank, world, &comm) != ICICLE_SUCCESS) return 11;
  size_t zlen = 0, wlen = 0;
  unsigned char* zkey = read_file(zkey_path, &zlen);
  unsigned char* wtns = read_file(wtns_path, &wlen);
  if (!zkey || !wtns) return 12;
  b200_zkey_cache* cache = NULL;
  /* precompute 0 = keep the window tables resident when they fit; every rank builds only the pieces b200_shard_plan gives it */
  if (b200_zkey_cache_create_sharded(zkey, zlen, 0, rank, world, &cache) != ICICLE_SUCCESS) return 13;
  uint32_t n_vars = 0;
  b200_zkey_cache_info(cache, &n_vars, NULL, NULL, NULL, NULL);
  if (wlen < (size_t)n_vars * 32) return 14;
  /* the witness values are the last n_vars * 32 bytes of the .wtns file (section 2, src/file_wrapper.rs:105-118) */
  const bn254_scalar_t* witness = (const bn254_scalar_t*)(wtns + wlen - (size_t)n_vars * 32);
  b200_groth16_proof proof;
  b200_prove_timings tm;
  eIcicleError e = b200_groth16_prove_sharded(cache, comm, witness, n_vars, NULL, NULL, rank == 0 ? &proof : NULL, &tm);
  if (e == ICICLE_SUCCESS && rank == 0) {
    static char json[4096];
    size_t n = b200_proof_to_json(&proof, json, sizeof json);
    FILE* f = fopen(proof_path, "w");
    if (!f || !n || fwrite(json, 1, n, f) != n) e = ICICLE_COPY_FAILED;
    if (f) fclose(f);
    printf("proof over %d GPUs in %.2f ms (device time of rank 0)\n", world, tm.total_ms);
  }
  b200_zkey_cache_destroy(cache);
  b200_comm_destroy(comm);
  free(zkey);
  free(wtns);
  return e == ICICLE_SUCCESS ? 0 : 15;
}

int main(int argc, char** argv)
{
  if (argc < 5) {
    fprintf(stderr, "usage: %s circuit.zkey witness.wtns proof.json n_gpus\n", argv[0]);
    return 2;
  }
  const int world = atoi(argv[4]);
  if (world < 1 || world > 64) return 2;
  /* rank 0 (this process) draws the token and hands it to its children through one pipe each */
  unsigned char token[128];
  if (b200_comm_unique_id(token) != ICICLE_SUCCESS) {
    fprintf(stderr, "no NCCL for the library (libnccl.so.2 not found)\n");
    return 3;
  }
  pid_t kids[64];
  for (int r = 1; r < world; ++r) {
    int fd[2];
    if (pipe(fd) != 0) return 4;
    kids[r] = fork();
    if (kids[r] == 0) {
      unsigned char t[128];
      close(fd[1]);
      if (read(fd[0], t, sizeof t) != (ssize_t)sizeof t) _exit(5);
      close(fd[0]);
      _exit(run_rank(r, world, t, argv[1], argv[2], argv[3]));
    }
    close(fd[0]);
    if (write(fd[1], token, sizeof token) != (ssize_t)sizeof token) return 6;
    close(fd[1]);
  }
  int rc = run_rank(0, world, token, argv[1], argv[2], argv[3]);
  for (int r = 1; r < world; ++r) {
    int st = 0;
    waitpid(kids[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 20 + r;
  }
  return rc;
}
