#!/bin/bash
# Build the experimental library variants measured by tools/evidence_r02.sh next to the default library (which is never
# built with experimental flags).  Run HERE before the gpurun call: the .so files are git-ignored and travel with the
# snapshot.  Each is a full build of the library (about a minute); they touch only k_seg_all's D <= 128 instantiation
# (what config 4 runs: 52 registers, 4 blocks of 256 threads per SM, 4.4 TB/s, long-scoreboard bound).
#   segmb6 / segmb8   6 / 8 resident blocks per SM (40 / 32 registers, 16 / 24 bytes spilled)
#   segpf             L2 prefetch of a window's staged contributions and table rows before the serial walk over its rows
#   segpf8            both
set -e
cd "$(dirname "$0")/.."
build() { DAISY_LIB_VARIANT=$1 DAISY_NVCC_EXTRA="$2" python -m recommend_lib_b200.build; }
build segmb6 "-DDAISY_SEG_MIN_BLOCKS_V1=6"
build segmb8 "-DDAISY_SEG_MIN_BLOCKS_V1=8"
build segpf "-DDAISY_SEG_PREFETCH=1"
build segpf8 "-DDAISY_SEG_PREFETCH=1 -DDAISY_SEG_MIN_BLOCKS_V1=8"
