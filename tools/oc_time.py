import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
from jatts_b200 import _lib
dev = torch.device("cuda", 0)
model, voc, cfg_fs2, cfg_hg = bench.build_models(dev)
from oracle import recipes
mels = [recipes.make_mel(305, i).to(dev) for i in range(64)]
for _ in range(3): y = voc.decode_batch(mels)
torch.cuda.synchronize()
ms = 0.0; n = 20
for _ in range(n):
    _lib.profile_begin(); y = voc.decode_batch(mels); c = _lib.profile_end_classes(); ms += c["output_conv"][0]
print("output_conv ms", ms / n, "tile", os.environ.get("JATTS_B200_OCT_TILE", "1024"), float(y[0].abs().mean()))
