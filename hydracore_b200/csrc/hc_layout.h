// hc_layout.h — binary layout facts of the blobs the layer receives from RenderDriverRTE.
// These are interface constants of the reference (Ray-Tracing-Systems/HydraCore); each block cites where the reference
// defines them.  tests/test_layout.py pins every HC_* value against the reference's own headers (tests/golden/ref_consts.json,
// dumped from oracle/_ref) so a drift is caught on CPU.  One `#define NAME value` per line: hydracore_b200/layout.py parses this.
#ifndef HC_LAYOUT_H
#define HC_LAYOUT_H

// ---- EngineGlobals header, byte offsets (hydra_drv/cfetch.h:21-81)
#define HC_EG_mProj 0
#define HC_EG_mWorldView 64
#define HC_EG_mProjInverse 128
#define HC_EG_mWorldViewInverse 192
#define HC_EG_varsI 256
#define HC_EG_varsF 512
#define HC_EG_rmQMC 768
#define HC_EG_camForward 832
#define HC_EG_imagePlaneDist 868
#define HC_EG_texturesTableOffset 872
#define HC_EG_materialsTableOffset 876
#define HC_EG_pdfTableTableOffset 880
#define HC_EG_geometryTableOffset 884
#define HC_EG_texturesAuxTableOffset 888
#define HC_EG_texturesTableSize 892
#define HC_EG_materialsTableSize 896
#define HC_EG_pdfTableTableSize 900
#define HC_EG_geometryTableSize 904
#define HC_EG_texturesAuxTableSize 908
#define HC_EG_floatArraysOffset 912
#define HC_EG_floatsArraysSize 916
#define HC_EG_lightSelectorTableOffsetRev 920
#define HC_EG_lightSelectorTableSizeRev 924
#define HC_EG_lightSelectorTableOffsetFwd 928
#define HC_EG_lightSelectorTableSizeFwd 932
#define HC_EG_g_flags 936
#define HC_EG_skyLightId 940
#define HC_EG_lightsOffset 944
#define HC_EG_lightsSize 948
#define HC_EG_lightsNum 952
#define HC_EG_sunNumber 968
#define HC_EG_suns 972
#define HC_EG_m_allTablesAreReady 5068
#define HC_EG_m_essGgx2017Table 5072
#define HC_EG_m_essTranspTable 13264
#define HC_EG_sizeof 537552
#define HC_EG_HEAD_BYTES 972

// ---- integer render variables varsI[] (hydra_drv/cglobals.h:440-490)
#define HC_HRT_ENABLE_DOF 0
#define HC_HRT_QMC_VARIANT 1
#define HC_HRT_TRACE_DEPTH 9
#define HC_HRT_DIFFUSE_TRACE_DEPTH 13
#define HC_HRT_RENDER_LAYER 19
#define HC_HRT_MMLT_FIRST_BOUNCE 34
#define HC_HRT_SHADOW_MATTE_BACK 35
#define HC_HRT_SHADOW_MATTE_BACK_MODE 41

// ---- float render variables varsF[] (hydra_drv/cglobals.h:492-537)
#define HC_HRT_DOF_LENS_RADIUS 0
#define HC_HRT_DOF_FOCAL_PLANE_DIST 1
#define HC_HRT_TILT_ROT_X 2
#define HC_HRT_TILT_ROT_Y 4
#define HC_HRT_IMAGE_GAMMA 6
#define HC_HRT_TEXINPUT_GAMMA 7
#define HC_HRT_ENV_COLOR_X 8
#define HC_HRT_ENV_COLOR_Y 9
#define HC_HRT_ENV_COLOR_Z 10
#define HC_HRT_CAM_FOV 14
#define HC_HRT_PATH_TRACE_ERROR 15
#define HC_HRT_PATH_TRACE_CLAMPING 16
#define HC_HRT_FOV_X 23
#define HC_HRT_FOV_Y 24
#define HC_HRT_WIDTH_F 25
#define HC_HRT_HEIGHT_F 26
#define HC_HRT_ABLOW_SCALE_X 29
#define HC_HRT_ABLOW_SCALE_Y 30

// ---- g_flags bits (hydra_drv/cglobals.h:405-436)
#define HC_HRT_COMPUTE_SHADOWS 1
#define HC_HRT_DISABLE_SHADING 2
#define HC_HRT_DIRECT_LIGHT_MODE 4
#define HC_HRT_UNIFIED_IMAGE_SAMPLING 8
#define HC_HRT_USE_MIS 32
#define HC_HRT_FORWARD_TRACING 256
#define HC_HRT_3WAY_MIS_WEIGHTS 1024
#define HC_HRT_ENABLE_MMLT 16384
#define HC_HRT_INDIRECT_LIGHT_MODE 65536
#define HC_HRT_STUPID_PT_MODE 524288
#define HC_HRT_ENABLE_PT_CAUSTICS 134217728

// ---- QMC remap slots rmQMC[] (hydra_drv/cglobals.h:81-95) and table shape (crandom.h:220-222)
#define HC_QMC_VAR_SCR_X 0
#define HC_QMC_VAR_SCR_Y 1
#define HC_QMC_VAR_DOF_X 2
#define HC_QMC_VAR_DOF_Y 3
#define HC_QMC_VAR_SRC_A 4
#define HC_QMC_VAR_MAT_L 5
#define HC_QMC_VAR_MAT_0 6
#define HC_QMC_VAR_MAT_1 7
#define HC_QMC_VAR_LGT_N 8
#define HC_QMC_VAR_LGT_0 9
#define HC_QMC_VAR_LGT_1 10
#define HC_QMC_VAR_LGT_2 11
#define HC_QRNG_DIMENSIONS_K 11
#define HC_QRNG_RESOLUTION_K 31
#define HC_MMLT_FLOATS_PER_BOUNCE 10
#define HC_MMLT_FLOATS_PER_SAMPLE 3
#define HC_MMLT_FLOATS_PER_MLAYER 7

// ---- ray flags word (hydra_drv/cglobals.h:1323-1365)
#define HC_RAY_EVENT_S 1
#define HC_RAY_EVENT_D 2
#define HC_RAY_EVENT_G 4
#define HC_RAY_EVENT_T 8
#define HC_RAY_EVENT_V 16
#define HC_RAY_EVENT_TOUT 32
#define HC_RAY_EVENT_TNINGLASS 64
#define HC_RAY_GRAMMAR_DIRECT_LIGHT 64
#define HC_RAY_GRAMMAR_OUT_OF_SCENE 128
#define HC_RAY_HIT_SURFACE_FROM_OTHER_SIDE 2048
#define HC_RAY_IS_DEAD 4096
#define HC_RAY_SHADE_FROM_OTHER_SIDE 8192
#define HC_RAY_SHADE_FROM_SKY_LIGHT 16384
#define HC_RAY_WILL_DIE_NEXT_BOUNCE 32768

// ---- traversal (hydra_drv/ctrace.h:576, 665-667)
#define HC_STACK_SIZE 80

#endif
