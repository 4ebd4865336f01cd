// spv2cpp — TEST INFRASTRUCTURE (builds the oracle/_ref reference arm; never part of the product path).
//
// Converts one of the reference's shipped SPIR-V modules (bin/data/Shaders/spirv/**.spv) into C++ with the
// reference's own vendored SPIRV-Cross C++ backend (dependencies/spirv-cross/spirv_cpp.cpp), so that the CPU
// reference arm executes exactly the arithmetic the reference ships. Output goes to stdout; the Makefile
// redirects it under oracle/_ref/gen/ (git-ignored — generated from reference files, never committed).
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <vector>

#include "spirv_cpp.hpp"

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: spv2cpp <module.spv>\n");
    return 2;
  }
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) {
    std::perror(argv[1]);
    return 1;
  }
  std::fseek(f, 0, SEEK_END);
  long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint32_t> words(static_cast<size_t>(bytes) / 4);
  size_t got = std::fread(words.data(), 4, words.size(), f);
  std::fclose(f);
  if (got != words.size() || words.empty()) {
    std::fprintf(stderr, "%s: short read\n", argv[1]);
    return 1;
  }
  spirv_cross::CompilerCPP compiler(std::move(words));
  std::cout << compiler.compile();
  return 0;
}
