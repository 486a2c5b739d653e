cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=0 FNSSL_TC_PAIR_MIN=1
echo "== debug 8 (mate does not report drained accumulators)"; FNSSL_TC_DEBUG=8 timeout 200 python tools/lstm_time.py full_in16_H128x2 full_in256_H128x2_add | grep -v "_B"
echo "== trace of the MATE CTA (debug 16)"; FNSSL_TC_DEBUG=16 timeout 200 python tools/tc5_trace.py 2>&1 | head -17
