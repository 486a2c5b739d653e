cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=0 FNSSL_TC_PAIR_MIN=1
F="full_in16_H128x2 full_in256_H128x2_add"
for d in 0 1 2 3 4 7; do echo "== tc5 flat FNSSL_TC_DEBUG=$d"; FNSSL_TC_DEBUG=$d timeout 300 python tools/lstm_time.py $F | grep -v "_B"; done
