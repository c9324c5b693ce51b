GPET_DET_NOTAIL=1 python tools/kprof.py --source source.txt --reps 20 --flush 2>&1 | grep -i "k_detector\|k_front\|per frame"
GPET_DET_NOTAIL=1 python tools/bigframes_sweep.py --scales 1,4 2>&1 | grep -i "k_detector"
GPET_DET_NOTAIL=1 GPET_REFILL_MIN=1 python tools/kprof.py --source source.txt --reps 20 --flush 2>&1 | grep -i "k_detector"
GPET_DET_NOTAIL=1 GPET_REFILL_MIN=8 python tools/kprof.py --source source.txt --reps 20 --flush 2>&1 | grep -i "k_detector"
