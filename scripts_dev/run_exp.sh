D=lenslesspicam_b200/_dev
LPC_LIB_PATH=$D/n3.so python scripts_dev/chk_variant.py 2>&1 | tail -1
LPC_LIB_PATH=$D/n7.so python scripts_dev/chk_variant.py 2>&1 | tail -1
for v in n3 n7 n3 n7; do
echo "### $v"
LPC_LIB_PATH=$D/$v.so scripts_dev/ab.sh "X=1" | tail -1
done
