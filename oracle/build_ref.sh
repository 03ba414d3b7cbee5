#!/bin/bash
# Build the reference's own native component (deblock.cpp) from where it lies under /root/reference, against the
# stand-in tiffio.h of oracle/deblock_ref/, into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
# Nothing of the reference is copied into the repo.  Usage: bash oracle/build_ref.sh [reference root]
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/_ref"
g++ -O2 -std=c++11 -w -I "$HERE/deblock_ref" "$REF/deblock.cpp" "$HERE/deblock_ref/tiff_stub.cpp" -o "$HERE/_ref/deblock_ref"
echo "built $HERE/_ref/deblock_ref"
