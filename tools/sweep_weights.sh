# segment-cost weights (diag, off-diag per k-step; fixed per segment) of the generate-once sweep at the kin40k bench shape (one slab):
# best two main-kernel ms of 8
for w in 5,8,64 5,8,32 5,8,100 5,8,16 10,16,64 10,16,100 10,17,80 21,33,150 21,33,250 21,34,150 20,33,150 13,21,100; do printf "WEIGHTS=%s: " $w; SGP_SWEEP4_WEIGHTS=$w timeout 60 python tools/profile_sweep.py 10000 512 8 2>&1 | awk '{print $7}' | sort -n | head -2 | tr '\n' ' '; echo; done
