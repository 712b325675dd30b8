run() { echo "== $*"; env "$@" bash runb.sh; }
run PHW_MQ_CFG=0
run PHW_MQ_CFG=5
run PHW_MQ_CFG=1
run PHW_MQ_CFG=2
run PHW_MQ_CFG=4
