#!/bin/sh
# TEST-ONLY: compile the CUDA sources against the thread-per-CUDA-thread emulation (see cuda_emu.h).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../../ace_jl_b200/csrc
OBJ=$HERE/_obj
mkdir -p "$OBJ"
pids=""
for u in "$SRC"/*.cu; do
    o="$OBJ/$(basename "$u" .cu).o"
    /usr/bin/g++ -std=c++17 -O1 -g -fPIC -DACEB200_EMU -Wno-unknown-pragmas -I"$HERE" -I"$SRC" -x c++ -c "$u" -o "$o" &
    pids="$pids $!"
done
for p in $pids; do wait $p; done
/usr/bin/g++ -shared -o "$HERE/libaceb200_emu.so" "$OBJ"/*.o -lpthread
