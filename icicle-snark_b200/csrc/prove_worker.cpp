// `prove` worker: the reference's stdin REPL (/root/reference/src/main.rs:39-186) over libicicle_b200.
// Same line protocol, defaults and sentinels: "prove --witness P --zkey P --proof P --public P --device CUDA",
// "exit"; prints COMMAND_EMPTY / COMMAND_EXIT / COMMAND_COMPLETED.  One process-wide cache (the CacheManager
// lives inside the library, keyed "{zkey}_{device}"), so the second proof for a zkey is warm.
// "verify --proof P --public P --vk P" runs the reference's verification equation (src/lib.rs:63-82,
// src/proof_helper.rs:319-372) with the library's host pairing - CPU code in the reference as well.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/icicle_b200.h"

static void print_help()
{
  printf("Usage: prove [OPTIONS]\n\n"
         "Options:\n"
         "--system <system>   Set the proof system (default: Groth16)\n"
         "--witness <path>    Path to the witness file\n"
         "--zkey <path>       Path to the zkey file\n"
         "--proof <path>      Path to the proof output file\n"
         "--public <path>     Path to the public output file\n"
         "--device <device>   Set the computation device (default: CUDA)\n"
         "--help              Show this message and exit\n");
}

static std::string lower(std::string s)
{
  for (auto& c : s)
    c = (char)tolower((unsigned char)c);
  return s;
}

int main()
{
  std::string line;
  for (;;) {
    printf("> ");
    fflush(stdout);
    if (!std::getline(std::cin, line)) break;
    std::istringstream is(line);
    std::vector<std::string> parts;
    for (std::string t; is >> t;)
      parts.push_back(t);
    if (parts.empty()) {
      printf("COMMAND_EMPTY\nCOMMAND_COMPLETED\n");
      fflush(stdout);
      continue;
    }
    if (lower(parts[0]) == "exit" && parts.size() == 1) {
      printf("COMMAND_EXIT\nCOMMAND_COMPLETED\n");
      fflush(stdout);
      break;
    }
    if (parts[0] == "prove") {
      std::string witness = "witness.wtns", zkey = "circuit_final.zkey", proof = "proof.json", pub = "public.json", device = "CUDA";
      bool ok = true;
      for (size_t i = 1; i < parts.size() && ok; ++i) {
        const std::string& a = parts[i];
        auto next = [&](std::string& dst) {
          if (i + 1 < parts.size())
            dst = parts[++i];
          else
            ok = false; // `parts.next()?` returns None
        };
        if (a == "--system") {
          if (i + 1 < parts.size() && lower(parts[++i]) != "groth16") {
            fprintf(stderr, "Unknown proof system: %s\n", parts[i].c_str());
            ok = false;
          }
        } else if (a == "--witness")
          next(witness);
        else if (a == "--zkey")
          next(zkey);
        else if (a == "--proof")
          next(proof);
        else if (a == "--public")
          next(pub);
        else if (a == "--device")
          next(device);
        else
          print_help();
      }
      if (!ok) {
        print_help();
        fflush(stdout);
        continue;
      }
      auto t0 = std::chrono::steady_clock::now();
      eIcicleError e = b200_groth16_prove_files(witness.c_str(), zkey.c_str(), proof.c_str(), pub.c_str(), device.c_str());
      double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (e != ICICLE_SUCCESS) {
        // the reference panics (unwrap) and takes the worker down; report and keep serving instead
        fprintf(stderr, "prove failed: eIcicleError %d%s\n", e, e == ICICLE_INVALID_DEVICE ? " (only --device CUDA is available: no CPU backend)" : "");
        printf("COMMAND_FAILED\n");
      } else {
        printf("proof took: %.6fs\n", s);
      }
      printf("COMMAND_COMPLETED\n");
      fflush(stdout);
      continue;
    }
    if (parts[0] == "verify") {
      std::string proof = "proof.json", pub = "public.json", vk = "verification_key.json";
      bool ok = true;
      for (size_t i = 1; i < parts.size() && ok; ++i) {
        const std::string& a = parts[i];
        auto next = [&](std::string& dst) {
          if (i + 1 < parts.size())
            dst = parts[++i];
          else
            ok = false;
        };
        if (a == "--system") {
          if (i + 1 < parts.size() && lower(parts[++i]) != "groth16") {
            fprintf(stderr, "Unknown proof system: %s\n", parts[i].c_str());
            ok = false;
          }
        } else if (a == "--proof")
          next(proof);
        else if (a == "--public")
          next(pub);
        else if (a == "--vk")
          next(vk);
        else
          print_help();
      }
      if (!ok) {
        print_help();
        fflush(stdout);
        continue;
      }
      int valid = 0;
      eIcicleError e = b200_groth16_verify_files(proof.c_str(), pub.c_str(), vk.c_str(), &valid);
      if (e != ICICLE_SUCCESS || !valid) {
        // the reference panics here (`assert!(pairing_result, "Verification failed")` / unwrap on a bad file)
        if (e != ICICLE_SUCCESS)
          fprintf(stderr, "verify failed: eIcicleError %d (unreadable or malformed input file)\n", e);
        else
          fprintf(stderr, "Verification failed\n");
        printf("COMMAND_FAILED\n");
      }
      printf("COMMAND_COMPLETED\n");
      fflush(stdout);
      continue;
    }
    print_help();
    fflush(stdout);
  }
  printf("Exiting CLI worker...\n");
  return 0;
}
