run() { echo "== $*"; bash runb.sh "$@"; }
run --phase 0 --extrap 2
run --phase 0
run
run --phase 0.5
run --debug-residuals
run --scheme SV
