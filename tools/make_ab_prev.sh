#!/bin/bash
# Snapshot the kernel sources of a commit (default HEAD) into tools/ab_prev/ so that tools/gpu_ab.sh
# can build them on the GPU box next to the working tree (the box has no .git).  Remove it afterwards.
rev=${1:-HEAD}
rm -rf tools/ab_prev
mkdir -p tools/ab_prev/mjpl_b200/csrc tools/ab_prev/include
for f in $(git ls-tree --name-only "$rev" mjpl_b200/csrc/); do git show "$rev:$f" > "tools/ab_prev/$f"; done
git show "$rev:include/mjpl_b200.h" > tools/ab_prev/include/mjpl_b200.h
echo "tools/ab_prev/ <- $rev"
