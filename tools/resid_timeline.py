import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 9), strict=True)
m.eval().cuda()
m.materialize.update(estimated_stfts=False, mask_per_speaker=False, spectrum=False, masks_b=False)
x = torch.from_numpy(np.tile(synth.make_mixtures(8, 64000, 1), (32, 1))).cuda()
for _ in range(3): m(x, {})
torch.cuda.synchronize()
os.environ["SEPTFA_FUSED_TL"] = "1"
m(x, {})
torch.cuda.synchronize()
