export AB_BASE="SPEECHT_B200_OVERLAP=0"
bash tools/gpu_ab.sh ab_overlap X= 2 3
