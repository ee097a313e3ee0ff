#!/bin/bash
# Builds an experimental libscfeat variant next to the product library (never loaded by default):
#   tools/build_variant.sh NAME "-DSCF_X=0 ..."   ->  tf-keras-speech-commands_b200/variants/libscfeat_NAME.so
# Select it at run time with SCFEAT_LIB=<path>.
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/tf-keras-speech-commands_b200/csrc
out=$root/tf-keras-speech-commands_b200/variants
mkdir -p $out /tmp/scf_$name
F="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$root/include -I$src $extra"
nvcc $F -c $src/scfeat_kernels.cu -o /tmp/scf_$name/k.o
nvcc $F -c $src/scfeat_host.cu -o /tmp/scf_$name/h.o
nvcc $F -c $src/scfeat_post.cu -o /tmp/scf_$name/p.o
nvcc $F -c $src/scfeat_ingest.cu -o /tmp/scf_$name/i.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libscfeat_$name.so /tmp/scf_$name/k.o /tmp/scf_$name/h.o /tmp/scf_$name/p.o /tmp/scf_$name/i.o -ldl -lpthread
echo $out/libscfeat_$name.so
