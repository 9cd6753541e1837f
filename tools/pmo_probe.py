"""Development probe: coeb_process_moving_object on one synthetic frame pair a few times (COEB_MOTION_TRACE=1 prints the host timeline;
under `ncu --metrics gpu__time_duration.sum` the launch list of one call)."""
import sys
sys.path[:0] = ['coeb-slam_b200/python']
from coeb_b200 import synth, motion
mo = motion.Motion()
prev, cur, _ = synth.make_motion_pair(0)
seq = len(sys.argv) > 2 and sys.argv[2] == "seq"   # the sequence form: frames fed round robin, the previous one resident on the device
call = mo.prepared_process_next([prev, cur]) if seq else mo.prepared_process(prev, cur)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    n = call()
print("ok", n)
