out=gpurun_out
python tools/bulk_debug.py ref
ODE_B200_LIB=$PWD/ode.jl_b200/libvar_bulk4.so python tools/bulk_debug.py bulk
python - <<'P'
import numpy as np
for n in (360000, 380000, 1<<20):
    a=np.load("/tmp/dbg_ref_%d.npy"%n); b=np.load("/tmp/dbg_bulk_%d.npy"%n)
    d=np.nonzero(a.view(np.uint64)!=b.view(np.uint64))[0]
    print(n, "differing points", d.size, d[:10], d[-10:] if d.size else "")
    if d.size:
        # runs of differing indices
        br=np.nonzero(np.diff(d)>1)[0]
        starts=np.concatenate([[d[0]], d[br+1]]); ends=np.concatenate([d[br],[d[-1]]])
        print(" runs:", [(int(s),int(e)) for s,e in zip(starts[:12],ends[:12])], "n_runs", len(starts))
P
rm -f $out/variants_c.log
bash tools/variant_run.sh variants_c ode.jl_b200/libode_b200.so ode.jl_b200/libvar_cnb.so > /dev/null 2>&1
cat $out/variants_c.log
ODE_B200_LIB=$PWD/ode.jl_b200/libvar_cnb.so python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for l in libode_b200 libvar_cnb; do ODE_B200_LIB=$PWD/ode.jl_b200/$l.so python tools/one_config.py c2 3 2>&1 | tail -1; done
