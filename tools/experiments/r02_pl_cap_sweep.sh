Q="--per-latent-keys --batch 4096 --steps 200 --no-cpu-baseline --sustain-seconds 0 --no-issue-rates --e2e-steps 1"
run() { env GSWM_LIB=$PWD/build_variants/$1.so GSWM_EMBED_PL_CTAS_PER_SM=$2 GSWM_EXTRACT_CTAS_PER_SM=$3 python bench.py $Q 2>/dev/null | python tools/benchq.py "plk[$1 embedcap=$2 extractcap=$3]"; }
run C2 0 0
run C2 4 1
run C2 3 1
run C2 2 2
run C3 4 1
run C3 3 1
run C4 4 1
run C4 3 1
run C6 4 1
run C6 3 1
run C6 2 1
