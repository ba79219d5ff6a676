import sys, statistics, numpy as np
sys.path.insert(0, "/root/repo")
import rmcv_b200 as rb
from rmcv_b200 import synth
W,H,B=1440,1080,64
raw=np.stack([synth.bgr_to_bayer(synth.make_frame(s,W,H,10),synth.BAYER_BG) for s in range(8)]*8)
c=rb.Context(max_width=W,max_height=H,max_batch=B)
d=c.device_buffer(raw.nbytes); m=c.device_buffer(B*H*W); d.upload(raw)
ms=[]
for i in range(10):
    c.timer_start(); c.bayer_extract_color_batch(d.ptr,W,H,B,synth.BAYER_BG,rb.CAMP_BLUE,80,m.ptr); t=c.timer_stop()
    if i>=3: ms.append(t)
t=statistics.median(ms)
print("bayer pixel stage: %.3f ms  %.0f GB/s (2 B/px)"%(t, B*H*W*2/(t*1e-3)/1e9))
