run() { echo "== $*"; env "$@" bash runb.sh --two-sweep; }
run PHW_P2_CFG=0
run PHW_P2_CFG=1
run PHW_P2_CFG=2
run PHW_P2_CFG=4
