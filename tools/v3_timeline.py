"""GPU box: in-kernel timeline of encoder v3 (libaft_b200_tl3.so, -DAFT_V3_TIMELINE): block 0, second sequence, layer 1;
slot 0 / 1 = compute warp 0 / 4 (streams 0 / 1), slot 2 = MMA issuer of stream 0.  Prints 'TL3 slot id clock' lines (stderr)."""
import os, sys
os.environ["AFT_ENCODER"] = "3"
os.environ["AFT_V3_TIMELINE_DUMP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import aft_oracle as O
from tests import util
sd = util.forti_weights(util.ada_weights())
m = util.make_model("forti", weights=sd, precision="bf16")
p, *_ = O.synthetic_batch(148 * 2, seed=3)      # 592 sequences: 4 per CTA
tp = torch.from_numpy(p).cuda()
with torch.no_grad():
    m(tp); torch.cuda.synchronize()
