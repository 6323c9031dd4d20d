D=lenslesspicam_b200/_dev
LPC_LIB_PATH=$D/base.so python scripts_dev/chk_variant.py 2>&1 | tail -1
LPC_LIB_PATH=$D/twp.so python scripts_dev/chk_variant.py 2>&1 | tail -1
for v in base twp base twp; do
echo "### $v"
LPC_LIB_PATH=$D/$v.so scripts_dev/ab.sh "X=1" | tail -1
LPC_LIB_PATH=$D/$v.so scripts_dev/ab_fista.sh "X=1" | tail -1
done
