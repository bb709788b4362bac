#!/bin/bash
# Diagnostic builds of libls_b200.so (selected at run time with LS_B200_LIB=<path>):
#   libls_prof.so     MMA-warp cycle accounting (LS_FUSED_TIMING=1 prints it)
#   libls_mc.so       cluster pairs with multicast weight stages (half the L2 reads, pair in lock step)
#   libls_nofetch.so  the producer signals stages without copying: pure MMA / epilogue timing, garbage results
#   libls_early.so    LS_LN1_EARLY=1: next block's LayerNorm-1 partial sums accumulated in the channel-mix epilogue
#   libls_mcearly.so  both
# (libls_mc.so and libls_early.so pass the 61 fused-path parity tests of tests/test_gpu_parity.py; A/B numbers in
#  profiles/r2_ab_power_cap.txt)
set -e
cd "$(dirname "$0")/../livelyspeaker_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
make -j8 libls_b200.so > /dev/null
build() {  # name, extra flags
  $NVCC $FLAGS $2 -c ls_fused.cu -o /tmp/ls_fused_$1.o
  $NVCC $ARCH -shared -o libls_$1.so $(ls *.o | grep -v '^ls_fused.o$' | grep -v 'umma\|bulk\|dsmem') /tmp/ls_fused_$1.o
}
build prof "-DLS_MMA_PROF=1" &
build mc "-DLS_MULTICAST=1 -DLS_MMA_PROF=1" &
build nofetch "-DLS_NOFETCH=1 -DLS_MMA_PROF=1" &
build early "-DLS_LN1_EARLY=1" &
build mcearly "-DLS_MULTICAST=1 -DLS_LN1_EARLY=1" &
wait
ls -la libls_*.so
