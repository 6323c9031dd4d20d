#!/bin/bash
# Quick build of the headline instantiations only (rows 3840, columns 2160) for A/B experiments:
#   scripts_dev/devbuild.sh OUT.so [-DFLAG=...]      then   LPC_LIB_PATH=OUT.so python bench.py --configs none ...
set -e
out=$1; shift
tmp=$(mktemp -d)
cat > $tmp/lpc_menu_extra.h <<'EOH'
#define LPC_ROW_MENU_EXTRA(X) X(3840, 15, 16, 16)
#define LPC_COL_MENU_EXTRA(X) X(2160, 15, 16, 9)
EOH
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -shared -o $out \
  $(dirname $0)/../lenslesspicam_b200/csrc/lpc_api.cu -L/usr/local/cuda/lib64 -lcufft -Xlinker -rpath=/usr/local/cuda/lib64 \
  -DLPC_MENU_ONLY_EXTRA -DLPC_HAVE_MENU_EXTRA_H -I$tmp "$@" 2>&1 | python $(dirname $0)/ptxas_summary.py
rm -rf $tmp
