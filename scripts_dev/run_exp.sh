D=lenslesspicam_b200/_dev
for v in l1 l5 l3 l7 l1 l5 l3 l7; do
echo "### $v"
LPC_LIB_PATH=$D/$v.so scripts_dev/ab.sh "X=1" | tail -1
done
