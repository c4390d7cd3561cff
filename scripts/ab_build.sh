# Experimental A/B build of libdkb.so (stride 4, 2 filter bits only; seconds instead of a minute).
# Usage: bash scripts/ab_build.sh NAME [-DFLAG ...]  ->  ab/libdkb_NAME.so ; run with DKB_LIBRARY=ab/libdkb_NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p ab
NAME=$1; shift
nvcc -shared -Xcompiler -fPIC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin g++ \
  -Xcompiler -pthread -DDKB_AB_BUILD "$@" -o ab/libdkb_$NAME.so denovo_kmer_b200/csrc/dkb_api.cu denovo_kmer_b200/csrc/dkb_host.cpp
echo ab/libdkb_$NAME.so
