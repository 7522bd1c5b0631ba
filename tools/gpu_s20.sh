export AB_BASE="SPEECHT_B200_LIB=speecht_b200/libspeecht_b200_base.so"
bash tools/gpu_ab.sh ab_dzprep X= 2
bash tools/gpu_ncu_set.sh ncu20 dzprep:ffa2_dz_prep:3 l10dgrad:tc_conv_kernel:64 l1dgrad:tc_conv_kernel:73
