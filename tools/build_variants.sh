#!/bin/bash
# Builds tuning variants of libuaes_b200.so into micro-aes_b200/lib_<name>/ (git-ignored; they travel to
# the GPU box with the snapshot).  tools/sweep_variants.py runs bench.py once per variant.
#   tools/build_variants.sh name1="-DFLAG=..." name2="..."
set -e
cd "$(dirname "$0")/../micro-aes_b200/csrc"
pids=()
for spec in "$@"; do
    name="${spec%%=*}"; flags="${spec#*=}"
    ( make LIBDIR=../lib_$name EXTRA="$flags" ../lib_$name/libuaes_b200.so ../lib_$name/libmicro_aes_128.so > /tmp/build_$name.log 2>&1 \
        && echo "built lib_$name ($flags)" || { echo "FAILED lib_$name"; tail -5 /tmp/build_$name.log; } ) &
    pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
