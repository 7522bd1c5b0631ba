export AB_BASE="SPEECHT_B200_BMN=0"
bash tools/gpu_ab.sh ab_bmn X= 2 3
bash tools/gpu_ncu_set.sh ncu19 packbwd:pack_bwd:2 packbwd9:pack_bwd:3 packboth:pack_filter_both:3 l9fwd:tc_conv_kernel:62
