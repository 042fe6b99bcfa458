timeout 600 python -m pytest tests/test_dsl.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/c8_pytest.log
tail -12 gpurun_out/c8_pytest.log
BK_DEBUG=1 python tools/gen_bench.py --brick 2>&1 | grep -v "k_star" | tee gpurun_out/c8_gen.log
BK_GEN_MINB=1 BK_DEBUG=1 python tools/gen_bench.py 2>&1 | grep -v "k_star" | tee gpurun_out/c8_gen_minb1.log
BK_GEN_MINB=2 BK_DEBUG=1 python tools/gen_bench.py 2>&1 | grep -v "k_star" | tee gpurun_out/c8_gen_minb2.log
