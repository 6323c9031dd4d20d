D=lenslesspicam_b200/_dev
LPC_LIB_PATH=$D/m4.so python scripts_dev/chk_variant.py 2>&1 | tail -1
LPC_LIB_PATH=$D/m3.so python scripts_dev/chk_variant.py 2>&1 | tail -1
for v in m4 m3 m4 m3; do
echo "### $v"
LPC_LIB_PATH=$D/$v.so scripts_dev/ab_fista.sh "X=1" | tail -1
done
