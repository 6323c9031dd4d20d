D=lenslesspicam_b200/_dev
export LPC_LIB_PATH=$D/cj.so
python scripts_dev/chk_variant.py 2>&1 | tail -1
LPC_CHAINS=3 LPC_CHAINS_JOINT_A=1 python scripts_dev/chk_variant.py 2>&1 | tail -1
scripts_dev/ab.sh "LPC_CHAINS=1" "LPC_CHAINS=3 LPC_CHAINS_JOINT_A=1" "LPC_CHAINS=1" "LPC_CHAINS=3 LPC_CHAINS_JOINT_A=1" "LPC_CHAINS=3 LPC_CHAINS_JOINT_A=1 LPC_ALTERNATE=0" "LPC_CHAINS=2 LPC_CHAINS_JOINT_A=1"
