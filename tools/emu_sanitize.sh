#!/bin/bash
# Run the CUDA kernels' per-thread code (host lane emulator) under AddressSanitizer + UBSan: the stand-in for
# compute-sanitizer, which is closed on the GPU pool.  Usage: bash tools/emu_sanitize.sh
set -e
cd "$(dirname "$0")/.."
cp tests/emu/libb2g_emu.so /tmp/libb2g_emu_backup.so 2>/dev/null || true
g++ -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -fsanitize=address,undefined -fno-omit-frame-pointer -Iinclude -Iisaacgymenv_b200/csrc \
    -x c++ tests/emu/emu.cpp -o tests/emu/libb2g_emu.so -lpthread
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
    python -m pytest tests/test_kernels_emu.py -x -q -p no:cacheprovider
# second pass: automatic variables start from a poison pattern -> any read of a link slot / scratch field that was never written
# (the rolled long-chain loops do not store unused slots) would surface as a parity failure
g++ -O1 -std=c++17 -fPIC -shared -ffp-contract=off -ftrivial-auto-var-init=pattern -Iinclude -Iisaacgymenv_b200/csrc \
    -x c++ tests/emu/emu.cpp -o tests/emu/libb2g_emu.so -lpthread
python -m pytest tests/test_kernels_emu.py -x -q -p no:cacheprovider
rm -f tests/emu/libb2g_emu.so
[ -f /tmp/libb2g_emu_backup.so ] && cp /tmp/libb2g_emu_backup.so tests/emu/libb2g_emu.so && touch tests/emu/libb2g_emu.so
