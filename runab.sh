run() { echo "== $* $EXTRA"; env "$@" bash runb.sh $EXTRA; }
EXTRA="--one-sweep" run PHW_DEGREE_SORT=0
EXTRA="--one-sweep" run PHW_DEGREE_SORT=1
EXTRA="" run PHW_DEGREE_SORT=0
EXTRA="" run PHW_DEGREE_SORT=1
