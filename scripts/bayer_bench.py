import sys, statistics, numpy as np
sys.path.insert(0, "/root/repo")
import rmcv_b200 as rb
from rmcv_b200 import synth
W,H,B=1440,1080,64
raw=np.stack([synth.bgr_to_bayer(synth.make_frame(s,W,H,10),synth.BAYER_BG) for s in range(8)]*8)
c=rb.Context(max_width=W,max_height=H,max_batch=B)
d=c.device_buffer(raw.nbytes); m=c.device_buffer(B*H*W); d.upload(raw)
ms=[]
REP=20   # launches per timed window: the event pair costs ~15 us of host enqueue time, a launch ~45 us
for i in range(10):
    c.timer_start()
    for _ in range(REP): c.bayer_extract_color_batch(d.ptr,W,H,B,synth.BAYER_BG,rb.CAMP_BLUE,80,m.ptr)
    t=c.timer_stop()/REP
    if i>=3: ms.append(t)
t=statistics.median(ms)
print("bayer pixel stage: %.4f ms per launch (avg of %d back-to-back launches)  %.0f GB/s (2 B/px)"%(t, REP, B*H*W*2/(t*1e-3)/1e9))
