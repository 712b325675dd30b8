run() { echo "== $*"; bash runb.sh "$@" --debug-residuals; }
run --extrap-q 4
run --extrap-q 5
run --extrap 4 --extrap-q 5
run --extrap 5 --extrap-q 5
