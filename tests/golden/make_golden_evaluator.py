"""Golden of the reference's OWN evaluation loop (row N4): ``ModelEvaluator.get_test_stats`` (src/main/trainer.py:311-347)
over ``get_test_dataloaders`` (src/data/dataset.py:193-262) on synthetic .mat test sets, with the seed-0 AdaFortiTran weights
of weights_ada_seed0.npz.  The test sets are written by the same seeded writer the GPU test uses
(tests/test_next_rows.py::_write_test_sets), so the test can rebuild the files byte for byte on the GPU box.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_evaluator.py     (needs /root/reference; writes golden_evaluator.npz)

matplotlib / prettytable (imported by the reference's src.utils, unused here) are not installed: both are stubbed.
"""
import logging, os, sys, tempfile, types
from pathlib import Path
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
from tests.test_next_rows import _write_test_sets, EVAL_SETS      # noqa: E402  (ours: the seeded .mat writer)
for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
    del sys.modules[k]
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot", "prettytable"):
    m = types.ModuleType(name)
    if name == "prettytable":
        m.PrettyTable = object
    sys.modules.setdefault(name, m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from src.config.schemas import ModelConfig, SystemConfig          # noqa: E402  (the reference's)
from src.data.dataset import get_test_dataloaders                  # noqa: E402
from src.main.trainer import ModelEvaluator                        # noqa: E402
from src.models import AdaFortiTranEstimator                       # noqa: E402
import src.models as ref_models                                    # noqa: E402
assert ref_models.__file__.startswith(REF), ref_models.__file__

sysc = SystemConfig(ofdm={"num_scs": 120, "num_symbols": 14}, pilot={"num_scs": 12, "num_symbols": 2})
modc = ModelConfig(model_type="adafortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4, activation="gelu",
                   dropout=0.1, max_seq_len=512, pos_encoding_type="learnable", channel_adaptivity_hidden_sizes=[7, 42, 560],
                   adaptive_token_length=6)
model = AdaFortiTranEstimator(sysc, modc).eval()
model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(HERE, "weights_ada_seed0.npz")).items()})
tmp = Path(tempfile.mkdtemp())
_write_test_sets(tmp, EVAL_SETS)
BATCH = 3
loaders = get_test_dataloaders(tmp, sysc.pilot, BATCH)
stats = ModelEvaluator(model, torch.device("cpu"), logging.getLogger("golden")).get_test_stats(loaders, torch.nn.MSELoss())
print(stats)
np.savez(os.path.join(HERE, "golden_evaluator.npz"), keys=np.array(list(stats.keys()), dtype=np.int64),
         mse_db=np.array(list(stats.values()), dtype=np.float64), batch_size=np.int64(BATCH))
