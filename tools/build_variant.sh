#!/bin/bash
# Builds an experimental libscfeat variant next to the product library (never loaded by default):
#   tools/build_variant.sh NAME "-DSCF_X=0 ..." [host]   ->  tf-keras-speech-commands_b200/variants/libscfeat_NAME.so
# Only the kernel file is recompiled (with -DSCF_VARIANT_BUILD: just the params.json fast-path kernels, seconds
# instead of a minute); the host objects come from the product build (run `make -C csrc` first).
# Select the variant at run time with SCFEAT_LIB=<path>.
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/tf-keras-speech-commands_b200/csrc
out=$root/tf-keras-speech-commands_b200/variants
mkdir -p $out /tmp/scf_$name
F="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$root/include -I$src -DSCF_VARIANT_BUILD $extra"
nvcc $F -Xptxas -v -c $src/scfeat_kernels.cu -o /tmp/scf_$name/k.o 2> /tmp/scf_$name/ptxas.log || (cat /tmp/scf_$name/ptxas.log; false)
hosto=$src/scfeat_host.o
if [ "$3" = "host" ]; then      # the switch also changes shared host/device geometry: rebuild the host side with it
    nvcc $F -c $src/scfeat_host.cu -o /tmp/scf_$name/h.o
    hosto=/tmp/scf_$name/h.o
fi
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libscfeat_$name.so /tmp/scf_$name/k.o $hosto $src/scfeat_post.o $src/scfeat_ingest.o -ldl -lpthread
grep -E "registers|spill" /tmp/scf_$name/ptxas.log | grep -v " 0 bytes spill" | sort | uniq -c | tr '\n' ';'
echo " -> $out/libscfeat_$name.so"
