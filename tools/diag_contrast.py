import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import signals, librosa_b200 as lb
from oracle import ref_np as O
y=signals.make("A",(9000,),seed=len("contrast_default_A"),sr=22050)
S=np.abs(O.stft(y)).astype(np.float32)
got=lb.feature.spectral_contrast(S=S,sr=22050,linear=True)
ref=O.spectral_contrast(S=S,sr=22050,linear=True)
d=np.abs(got-ref)
bad=np.argwhere(d>1e-4*np.abs(ref)+1e-7)
print("bad cells",len(bad))
for b,t in bad[:12]: print(b,t,got[b,t],ref[b,t])
