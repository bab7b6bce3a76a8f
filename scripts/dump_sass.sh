#!/bin/bash
# scripts/dump_sass.sh -- the SASS listings DESIGN.md's instruction counts refer to, from the objects
# `make -C r-star-tree_b200/csrc` built (sm_100a), into profiles/.
#   usage: scripts/dump_sass.sh [tag]     (tag defaults to r2)
set -e
cd "$(dirname "$0")/.."
tag=${1:-r2}
L=r-star-tree_b200/csrc
dump() { # object, mangled name, output, title
  {
    echo "// $4"
    echo "// $(cuobjdump -res-usage "$1" 2>/dev/null | grep -A1 "$2" | tail -1 | sed 's/^ *//')"
    echo "// cuobjdump -sass -fun $2 $1   (encodings stripped)"
    cuobjdump -sass -fun "$2" "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/\s+$//' | awk '{$1=$1};1'
  } > "$3"
  echo "$3: $(wc -l < "$3") lines"
}
dump $L/rt_box_f32.o _ZN3rtb19box_traverse_kernelIfLi3ELi8ELb0ELb0ELj2ELb1ELb0EEEvNS_9BoxLaunchE \
     profiles/${tag}_box_traverse_f32_3d_w8_stash_quant.sass \
     "rtb::box_traverse_kernel<float, 3, 8, point keys, PASS_STASH, QUANT> -- the headline kernel (BASELINE config 2)"
dump $L/rt_box_f32.o _ZN3rtb19box_traverse_kernelIfLi8ELi16ELb0ELb0ELj2ELb1ELb0EEEvNS_9BoxLaunchE \
     profiles/${tag}_box_traverse_f32_8d_w16_stash_quant.sass \
     "rtb::box_traverse_kernel<float, 8, 16, point keys, PASS_STASH, QUANT> -- BASELINE config 4's kernel"
dump $L/rt_knn.o _ZN3rtb19knn_traverse_kernelILi3ELi8ELb0ELb0EEEvNS_9KnnLaunchE \
     profiles/${tag}_knn_traverse_3d_w8.sass \
     "rtb::knn_traverse_kernel<3, 8, point keys>"
dump $L/rt_ray.o _ZN3rtb19ray_traverse_kernelILi8ELb1ELj3EEEvNS_9RayLaunchE \
     profiles/${tag}_ray_traverse_w8_boxkeys_closest.sass \
     "rtb::ray_traverse_kernel<8, box keys, RAY_CLOSEST>"
