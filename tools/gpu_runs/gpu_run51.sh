for K in 1 2 4; do BROADCAST_B200_STRIP_CHAINS=$K timeout 300 python tools/jac_probe.py 2048x512 4096x1024 2>&1 | cut -c1-140 | sed "s/^/chains=$K /"; done
