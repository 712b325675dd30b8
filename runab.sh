for m in 0 1 2; do echo "== PHW_ELL_BAR=$m"; PHW_ELL_BAR=$m bash profiles/runb.sh 2>&1 | tail -1; done
