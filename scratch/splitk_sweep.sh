# C5 training step time against the split-K rule of the layer launches (k-ranges per tile <= smax, >= tmin k-chunks per range)
for cfg in ${SWEEP:-"4 4" "4 2" "4 1" "2 2" "2 1"}; do set -- $cfg; LINNA_TG_MAX_SPLITK=$1 LINNA_TG_SPLITK_MIN_CHUNKS=$2 python scratch/tg_one.py 2>&1 | tail -1 | sed "s/^/smax=$1 tmin=$2: /"; done
