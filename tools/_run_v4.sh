for cfg in "gv4 x" "r4 x" "r2 x" "r2 12,1"; do set -- $cfg; echo "== $1 force $2"; if [ "$2" != x ]; then export GGQ_PLAN_FORCE=$2; else unset GGQ_PLAN_FORCE; fi; GGQ_VARIANT=$1 timeout 120 python tools/probe_variant.py 2>&1; done
unset GGQ_PLAN_FORCE
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
