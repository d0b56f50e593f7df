#!/bin/bash
# tools/build_variant.sh <name> <nvcc -D flags...>: a kernel-variant build of both libraries under variants/<name>/ (git-ignored; travels to the GPU box)
name=$1; shift
make -s -j8 LIBDIR=variants/$name OBJDIR=build/obj_$name EXTRA="$*" variants/$name/libvanrijn_cuda.so variants/$name/libvanrijn_host.so 2>&1 | grep -E "error|Error" 
ls variants/$name/*.so >/dev/null && echo "built variants/$name ($*)"
