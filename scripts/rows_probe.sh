for r in 0 8 16 32 64 128; do echo "FWD_ROWS=$r"; RISP_FUSED_FWD_ROWS=$r PROBE_TAG=rows$r python scripts/probe_step.py bilinear 2>&1 | grep -E "A_gain|demosaic" ; done
