#!/bin/bash
# AddressSanitizer + UBSan over the host builds of the decoders on adversarial inputs (no GPU needed):
#  - the oracle (oracle/*.c): reference verdict vectors, golden LZ4 vectors, mutated .4mc / .4mz files
#  - the product's host-compilable sources (zstd_decode.h, lz4_parse.h -- the same code the kernels run):
#    the same vectors with exact-size source buffers, plus mutated frames of the reference's .4mz files
set -e
cd "$(dirname "$0")/.."
D=/tmp/fourmc_asan; mkdir -p $D
F="-O1 -g -fsanitize=address,undefined -shared -fPIC"
gcc $F -std=c99 -o $D/liboracle_asan.so oracle/fourmc_oracle.c oracle/zstd_oracle.c
g++ $F -std=c++17 -o $D/zstd_shim_asan.so tests/native/zstd_shim.cpp
g++ $F -std=c++17 -o $D/parse_shim_asan.so tests/native/parse_shim.cpp
export LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0
python tools/asan_oracle_cases.py
python tools/asan_product_cases.py
g++ $F -std=c++17 -o $D/zenc_emul_asan.so tests/native/zenc_emul.cpp
python tools/asan_encoder_cases.py
g++ $F -std=c++17 -D_GLIBCXX_SANITIZE_VECTOR -o $D/enc_emul_asan.so tests/native/enc_emul.cpp
python tools/asan_lz4_encoder_cases.py
