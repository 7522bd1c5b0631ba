export AB_BASE="SPEECHT_B200_LIB=speecht_b200/libspeecht_b200_base.so"
bash tools/gpu_ab.sh ab_epw16 X= 2 3
python tools/ctc_bench.py > gpurun_out/ab_epw16/ctc_bench.txt 2>&1; tail -3 gpurun_out/ab_epw16/ctc_bench.txt
