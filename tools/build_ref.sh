#!/bin/bash
# builds libsgp of a git ref (default HEAD) for same-box A/B timing: tools/build_ref.sh <name> [ref]
# -> gaussianprocessnode_b200/libsgp_<name>.so ; use it with SGP_LIB_PATH=...
set -e
name=$1; ref=${2:-HEAD}
root="$(cd "$(dirname "$0")/.." && pwd)"
tmp=$(mktemp -d)
git -C "$root" archive "$ref" gaussianprocessnode_b200/csrc include | tar -x -C "$tmp"
cd "$tmp/gaussianprocessnode_b200"
objs=""
for f in csrc/*.cu; do
  o="$tmp/$(basename "$f" .cu).o"; objs="$objs $o"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -c "$f" -o "$o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/gaussianprocessnode_b200/libsgp_$name.so" $objs -lcudart -ldl
rm -rf "$tmp"
echo "libsgp_$name.so ($ref)"
