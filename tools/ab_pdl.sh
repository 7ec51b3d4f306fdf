for rep in 1 2; do for pdl in 0 1; do export GGS_B200_PDL=$pdl; echo "== PDL=$pdl rep $rep";
TQDM_DISABLE=1 python examples/run_ga_synthetic.py --side 256 --splats 512 --pop 32 --generations 5000
TQDM_DISABLE=1 python examples/run_ga_synthetic.py --side 128 --splats 100 --pop 32 --generations 5000
python tools/time_sa.py 2>&1 | grep "device engine"
done; done
