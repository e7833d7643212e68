#!/bin/sh
# TEST-ONLY: compile the CUDA sources against the thread-per-CUDA-thread emulation (see cuda_emu.h).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../../ace_jl_b200/csrc
/usr/bin/g++ -std=c++17 -O1 -g -fPIC -shared -DACEB200_EMU -Wno-unknown-pragmas -I"$HERE" -I"$SRC" \
    -x c++ "$SRC/aceb200.cu" -o "$HERE/libaceb200_emu.so" -lpthread
