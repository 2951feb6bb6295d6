set -e
for f in examples/mcmc/mcmc_prob4a.py examples/mcmc/metrohast_norm1d.py examples/mcmc/gibbs_norm2d.py examples/dgei/dgei_norm1d_improved.py; do echo "== $f"; PYTHONPATH=. python $f 2>&1 | tail -4; done
echo "== mcmc_prob4a 4096 chains"; PYTHONPATH=. python examples/mcmc/mcmc_prob4a.py 4096 2>&1 | tail -4
echo "== mh_linreg"; PYTHONPATH=. python examples/mcmc/mh_linreg.py 100000 2048 2>&1 | tail -6
echo "== dgei 4096 grid N=1e5"; PYTHONPATH=. python examples/dgei/dgei_norm1d_improved.py 4096 100000 2>&1 | tail -3
echo "== omc_rs_sp_norm1d"; PYTHONPATH=. python examples/omc/omc_rs_sp_norm1d.py 2>&1 | tail -3
echo "== omc_rs_sp_norm1d 1e7 samples"; PYTHONPATH=. python examples/omc/omc_rs_sp_norm1d.py 10000000 1000 2>&1 | tail -3
