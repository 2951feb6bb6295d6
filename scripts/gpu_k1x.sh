for np in 11 13; do
PBX_NVCC_EXTRA="-DWS_BIGCTA_MAXD=0 -DWS_SMALL_NPROD=$np" python -m probayes_b200.build --force > /dev/null 2>&1
echo "NP=$np"; python scripts/bench_k1_dims.py 2>&1 | sed -n 2,4p
done
