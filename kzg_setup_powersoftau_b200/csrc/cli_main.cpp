// preprocess-kgz / preprocess-fastkgz: drop-in replacements of the reference's two
// binaries (/root/reference/src/bin/preprocess-kgz.rs, preprocess-fastkgz.rs).
// With no arguments they behave like the reference: cwd files `powersoftau` ->
// `powersoftau_uncompressed` -> `kzg_setup`, 2^21 powers, BLAKE2b digest check of the
// ceremony file.  The reference takes no options; the ones below exist because its
// constants are compile-time.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <string>

#include "../../include/ptau_b200.h"

#ifndef PTAU_CLI_VARIANT
#define PTAU_CLI_VARIANT PTAU_VARIANT_KGZ
#endif

static void usage(const char* argv0) {
  fprintf(stderr,
          "usage: %s [--dir D] [--log2-powers K] [--skip-digest | --expect-digest HEX] [--no-uncompressed]\n"
          "          [--gpus N] [--checks strict|reference] [--fsync]\n",
          argv0);
}

int main(int argc, char** argv) {
  std::string dir = ".";
  unsigned log2n = 21, flags = 0, checks = PTAU_CHECKS_STRICT;
  int gpus = 1;
  const char* digest = nullptr;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--dir" && i + 1 < argc) dir = argv[++i];
    else if (a == "--log2-powers" && i + 1 < argc) log2n = (unsigned)atoi(argv[++i]);
    else if (a == "--skip-digest") flags |= PTAU_FILE_SKIP_DIGEST;
    else if (a == "--expect-digest" && i + 1 < argc) digest = argv[++i];
    else if (a == "--no-uncompressed") flags |= PTAU_FILE_NO_UNCOMPRESSED;
    else if (a == "--fsync") flags |= PTAU_FILE_FSYNC;
    else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
    else if (a == "--checks" && i + 1 < argc) {
      std::string c = argv[++i];
      if (c == "strict") checks = PTAU_CHECKS_STRICT;
      else if (c == "reference") checks = PTAU_CHECKS_READ;
      else { usage(argv[0]); return 2; }
    } else { usage(argv[0]); return 2; }
  }
  ptau_ctx* ctx = nullptr;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int rc = ptau_create(&ctx, gpus, nullptr, 0);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (getenv("PTAU_TRACE"))
    fprintf(stderr, "[ptau trace] ptau_create (CUDA context)   %.1f ms\n", (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) / 1e6);
  if (rc != PTAU_OK) {
    fprintf(stderr, "ptau_create(%d GPUs): %s\n", gpus, ptau_strerror(rc));
    return 1;
  }
  std::string in = dir + "/powersoftau", unc = dir + "/powersoftau_uncompressed", out = dir + "/kzg_setup";
  if (!(flags & PTAU_FILE_SKIP_DIGEST)) printf("Checking existing powersoftau file...\n");
  printf("Started deserializing compressed Powers of Tau...\n");
  uint64_t bad_i = 0;
  int bad_k = 0, bad_s = -1;
  rc = ptau_preprocess_files(ctx, PTAU_CLI_VARIANT, in.c_str(), out.c_str(), unc.c_str(), log2n, digest, flags, checks, &bad_i,
                             &bad_k, &bad_s);
  if (rc > 0) {
    static const char* names[5] = {"tau_powers_g1", "tau_powers_g2", "alpha_tau_powers_g1", "beta_tau_powers_g1", "beta_g2"};
    fprintf(stderr, "InvalidData: %s at point %llu of %s\n", ptau_strerror(rc), (unsigned long long)bad_i,
            bad_s >= 0 && bad_s < 5 ? names[bad_s] : "?");
  } else if (rc == PTAU_ERR_SIZE) {
    fprintf(stderr, "The size of `powersoftau` should be %llu, so something isn't right.\n",
            (unsigned long long)ptau_response_size(1ull << log2n));
  } else if (rc == PTAU_ERR_EXISTS) {
    fprintf(stderr, "unable to create `powersoftau_uncompressed`\n");
  } else if (rc == PTAU_ERR_DIGEST) {
    fprintf(stderr, "failed validation (digest mismatch; the reference would now download, impossible offline)\n");
  } else if (rc != PTAU_OK) {
    fprintf(stderr, "error: %s (%s)\n", ptau_strerror(rc), ptau_last_error(ctx));
  } else {
    printf("Done serializing. KZG parameters are stored in kzg_setup\n");
  }
  // every output is closed and published (or unlinked) by now: leave without tearing the CUDA context and the
  // pinned slabs down page by page
  fflush(stdout);
  fflush(stderr);
  _exit(rc == PTAU_OK ? 0 : 1);
}
