for w in 0 1 2; do echo "== active warps $w (0 = all four)"; MPS_ROLLOUT_ACTIVE_WARPS=$w python profiles/tools/chain_cycles.py 2>&1 | tail -16; done
