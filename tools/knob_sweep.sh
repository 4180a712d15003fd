run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['config']['pcg_iterations'], round(d['roofline']['kernels']['vcycle']['ms'],3))"; }
run A=1
run APM_LCL_NU_L1=1
run APM_LCL_NU_L1=3
run APM_LCL_NU_DEEP=1
run APM_LCL_NU_DEEP=3
run APM_LCL_W_LO=2
run APM_LCL_W_HI=5
run APM_LCL_W_HI=6
run APM_LCL_W_LO=2 APM_LCL_W_HI=6
run APM_LCL_W_LO=4
run APM_LCL_W_LO=1 APM_LCL_W_HI=0
