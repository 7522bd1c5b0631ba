bash tools/gpu_ab.sh ab_wgrad_pair "SPEECHT_B200_WGRAD_PAIR=1" 2 3
bash tools/gpu_ncu_set.sh ncu17 combine:ffa2_combine:3 dzprep:ffa2_dz_prep:3 dxcomb:ffa2_dx_combine:3 dwcomb:ffa2_dw_combine:3 packffa2:pack_ffa2:3 packboth:pack_filter_both:3 adam:clip_adam:3 l10dgrad:tc_conv_kernel:63 l10fwd:tc_conv_kernel:63 ctcab:ctc_alpha_beta:3
