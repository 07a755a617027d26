#!/bin/bash
# An alternative build of libmfit_cuda.so for A/B runs on one box (not product code):
#   tools/build_variant.sh <name> [-DMFIT_WARPS=4 -DMFIT_MIN_CTAS=3 ...]   ->  gpurun_variants/libmfit_<name>.so
# Recompiles the translation units that see the kernel shape (api.cu, eval_builtin_ll.cu per likelihood kind) with the extra flags and links them
# with the other objects of the regular build (run `make -C lisp_mcmc_b200/csrc` first).  Use it with MFIT_LIB=<path>.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/lisp_mcmc_b200/csrc
out=$root/gpurun_variants
mkdir -p $out/$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
$NV "$@" -c -o $out/$name/api.o $src/api.cu &
$NV "$@" -DMFIT_EB_LL=0 -Xptxas -v -c -o $out/$name/eval_builtin_normal.o $src/eval_builtin_ll.cu 2> $out/$name/ptxas.log &
$NV "$@" -DMFIT_EB_LL=1 -c -o $out/$name/eval_builtin_cutoff.o $src/eval_builtin_ll.cu &
$NV "$@" -DMFIT_EB_LL=2 -c -o $out/$name/eval_builtin_poisson.o $src/eval_builtin_ll.cu &
wait
objs=""
for o in eval_builtin eval_perchain eval_perchain_reg_normal eval_perchain_reg_poisson eval_fit chain_kernels prep_kernels measure lower nvrtc_glue nccl_glue multi eval_src models_src; do objs="$objs $src/build/$o.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out/libmfit_$name.so $out/$name/api.o $out/$name/eval_builtin_normal.o $out/$name/eval_builtin_cutoff.o $out/$name/eval_builtin_poisson.o $objs -ldl -lpthread
grep -A2 "GaussLinLi0EEv13MfitMultiArgs" $out/$name/ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " -> $out/libmfit_$name.so"
