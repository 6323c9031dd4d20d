D=lenslesspicam_b200/_dev
export LPC_GD_ROWS_TMA=0
echo "#### v0"; LPC_LIB_PATH=$D/v0.so python scripts_dev/chk_variant.py 2>&1 | tail -1
for v in z1 p1 p2; do
echo "### $v"
LPC_LIB_PATH=$D/$v.so python scripts_dev/chk_variant.py 2>&1 | tail -1
LPC_LIB_PATH=$D/$v.so scripts_dev/ab_fista.sh "X=1" | tail -1
done
