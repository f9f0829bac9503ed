"""Drop-in checks with the reference's OWN caller code (1) and (2).

(1) `RunGAN.train_disc` (run_gun.py:339-398, the WGAN-GP critic loop with
its double backward and Adam steps) is imported unchanged from /root/reference and run twice on identical seeded inputs -
once on the reference's DiscV2 (in a subprocess, where `models` resolves to the reference package) and once on this
repo's DiscV2 (`models` first on sys.path, exactly the integration of INTEGRATION.md 1; kernels emulated on CPU, fp32).
The critic weights after the loop and the two returned running sums must agree.

(2) `evaluate.gather_results` (evaluate.py:101-116: the inference loop of the evaluator - `net(frames, regions, None)` then
`net.decoder.decode_tokens` per clip) is imported unchanged and run over a fake loader with the reference's CapGnnModel and
with ours, greedy and beam-3: the caption strings per video id must be identical.

(3) The whole `RunGAN(...).train()` (run_gun.py:20-320: model construction with the msr-vtt overrides, both Adams and
schedulers, GANLambdaHandler, per iteration G forward -> 2 critic steps -> G forward -> packed CE -> critic term -> backward
-> Adam, evaluation hook at the saving schedule) runs unchanged for one epoch of two batches on the reference's models and
on ours, from the same seeds (default initialisation: identical by construction, see tests/golden/init_order.json).  The
generator and critic weights after training must agree.  Environment patches, identical for both runs: dropout is held off
(`nn.Module.train` keeps eval mode; the RNG streams of custom dropout kernels cannot match torch's), `SummaryWriter` and
`evaluate.evaluate` are stand-ins (tensorboard files / java scorers), the working directory is a scratch directory.

(5) Edge shapes straight against the reference modules (not via the oracle): a batch of one clip, and num_obj = 4 where the
reference constructs no obj_embed and skips the region aggregation (layer.py:143,182-183): logits, nodes, attention weights
and every parameter gradient.

(4) The baseline trainer `Run(...).train()` (run_graph.py:19-200; what train.py runs: CapBaseline1 with the msr-vtt override
decode_hidden_size = 1300, an odd width for the GEMM / LSTM kernels) for two epochs of ten batches - including its every-10-
steps sample print through `decoder.decode_tokens`, the MultiStepLR milestone and the evaluation hook - same comparison.

Needs the reference checkout (only present in the build container): skipped elsewhere.  Stubs: `evaluate` (for run_gun: it would
import h5py / tables), `utils.data` and `cocoeval` (for evaluate.py: h5py / java), `seaborn`, `matplotlib.pyplot`,
`allennlp.common.checks`."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('DLSG_REFERENCE', '/root/reference')

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'run_gun.py')), reason='reference checkout not present')

WORKER = r'''
import os, sys, types, contextlib, io
ROOT, REF, which, out = sys.argv[1:5]
pkg = os.path.join(ROOT, 'd-lsg-video-caption_b200')
for name in ('allennlp', 'allennlp.common', 'allennlp.common.checks', 'seaborn', 'matplotlib', 'matplotlib.pyplot', 'evaluate'):
    sys.modules[name] = types.ModuleType(name)
sys.modules['allennlp.common.checks'].ConfigurationError = type('ConfigurationError', (Exception,), {})
for n in ('evaluate', 'convert_data_to_coco_scorer_format', 'gather_results', 'evaluate_multi_gpu'):
    setattr(sys.modules['evaluate'], n, None)
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
if which == 'reference':
    sys.path[:0] = [REF, pkg, ROOT]                 # `models` = the reference's package
else:
    sys.path[:0] = [pkg, ROOT, os.path.join(ROOT, 'tests'), REF]      # `models` = ours; run_gun / utils come from the reference
import numpy as np
import torch
torch.set_num_threads(4)
with contextlib.redirect_stdout(io.StringIO()):
    import run_gun                                   # the reference's trainer module, unmodified
import models
assert models.__file__.startswith(REF if which == 'reference' else pkg), models.__file__
assert run_gun.__file__.startswith(REF)
from dlsg import synth
if which == 'ours':
    from dlsg import ops, linalg as la
    from cpu_emul import CpuEmulBackend
    ops.set_backend(CpuEmulBackend())
    la.set_precision('fp32')
from models.model import DiscV2
args = synth.small_args(visual_hidden_size=1024, num_proposals=5, num_topk=5)
V, B, L = 37, 3, args.max_words
D = DiscV2(args, V)
synth.fill_state_dict(D, prefix='D.')
D.eval()
rs = np.random.RandomState(5)
_, _, caps, lens = synth.make_inputs(B, args, V, seed=14)
att_mask = synth.att_mask_from_captions(caps)
obj = torch.from_numpy(rs.standard_normal((B, 5, 1024)).astype(np.float32))
mot = torch.from_numpy(rs.standard_normal((B, 5, 1024)).astype(np.float32))
alpha = torch.softmax(torch.from_numpy(rs.standard_normal((B, L, 10)).astype(np.float32)), -1)
fake = torch.from_numpy(rs.standard_normal((B, L, V)).astype(np.float32))
trainer = run_gun.RunGAN.__new__(run_gun.RunGAN)     # no __init__: it would build loaders, tensorboard files, ...
trainer.device, trainer.multi_gpu, trainer.local_rank = torch.device('cpu'), False, 0
trainer.writer = types.SimpleNamespace(add_scalar=lambda *a, **k: None)
real = trainer.to_onehot(caps, V)                    # run_gun.py:449-453
opt_d = torch.optim.Adam(D.parameters(), lr=1.6e-4, betas=(0.5, 0.9))
torch.manual_seed(5)
loss_sum, wass = trainer.train_disc(real, fake, opt_d, D, 3, 1, 0, 10, 0.0, 0.0, obj_psl=obj, motion_psl=mot, pos_tag=None,
                                    att_mask=att_mask, alpha_all=alpha)
np.savez(out, loss_sum=np.float64(loss_sum), wass=np.float64(wass),
         **{'p.' + k: v.detach().numpy() for k, v in D.state_dict().items()})
'''


EVAL_WORKER = r'''
import os, sys, types, contextlib, io, json
ROOT, REF, which, out = sys.argv[1:5]
pkg = os.path.join(ROOT, 'd-lsg-video-caption_b200')
for name in ('allennlp', 'allennlp.common', 'allennlp.common.checks', 'cocoeval', 'utils.data'):
    sys.modules[name] = types.ModuleType(name)
sys.modules['allennlp.common.checks'].ConfigurationError = type('ConfigurationError', (Exception,), {})
sys.modules['cocoeval'].COCOScorer = sys.modules['cocoeval'].suppress_stdout_stderr = None
sys.modules['utils.data'].get_eval_loader = None
if which == 'reference':
    sys.path[:0] = [REF, pkg, ROOT]
else:
    sys.path[:0] = [pkg, ROOT, os.path.join(ROOT, 'tests'), REF]
import torch
torch.set_num_threads(4)
import evaluate                                      # the reference's evaluator module, unmodified
import models
assert models.__file__.startswith(REF if which == 'reference' else pkg), models.__file__
assert evaluate.__file__.startswith(REF)
from dlsg import synth
if which == 'ours':
    from dlsg import ops, linalg as la
    from cpu_emul import CpuEmulBackend
    ops.set_backend(CpuEmulBackend())
    la.set_precision('fp32')
from models.model import CapGnnModel
args = synth.small_args()
V, B = 37, 3
with contextlib.redirect_stdout(io.StringIO()):
    net = CapGnnModel(args, synth.Vocab(V))
synth.fill_state_dict(net)
net.eval()
loader = []
for b in range(2):
    fr, rg, _, _ = synth.make_inputs(B, args, V, seed=40 + b)
    loader.append((fr, rg, None, ['vid%d_%d' % (b, i) for i in range(B)]))
res = {}
with torch.no_grad():
    for beam in (1, 3):
        net.update_beam_size(beam)
        captions, _ = evaluate.gather_results(net, args, loader)
        res['beam%d' % beam] = dict(captions)
json.dump(res, open(out, 'w'))
'''


TRAIN_WORKER = r'''
import os, sys, types, contextlib, io, random
ROOT, REF, which, out = sys.argv[1:5]
os.chdir(os.path.dirname(out))
pkg = os.path.join(ROOT, 'd-lsg-video-caption_b200')
for name in ('allennlp', 'allennlp.common', 'allennlp.common.checks', 'seaborn', 'matplotlib', 'matplotlib.pyplot', 'evaluate'):
    sys.modules[name] = types.ModuleType(name)
sys.modules['allennlp.common.checks'].ConfigurationError = type('ConfigurationError', (Exception,), {})
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


def fake_evaluate(net, opt, loader, reference, multi_modal=False, multi_gpu=False):
    res = {}
    for frames, regions, _, vids in loader:                      # what evaluate.py:56-98 does, minus the java scorers
        outputs = net(frames, regions[:, :, :opt.num_obj, :], None)[0]
        for tokens, vid in zip(outputs, vids):
            res[vid] = net.decoder.decode_tokens(tokens.data)
    return {'Bleu_4': 0.1, 'METEOR': 0.1, 'CIDEr': 0.1, 'ROUGE_L': 0.1}, res, [], 0.0


ev = sys.modules['evaluate']
ev.evaluate = fake_evaluate
ev.convert_data_to_coco_scorer_format = ev.gather_results = ev.evaluate_multi_gpu = None
if which == 'reference':
    sys.path[:0] = [REF, pkg, ROOT]
else:
    sys.path[:0] = [pkg, ROOT, os.path.join(ROOT, 'tests'), REF]
import numpy as np
import torch
import torch.nn as nn
torch.set_num_threads(4)
_train, _init = nn.Module.train, nn.Module.__init__
nn.Module.train = lambda self, mode=True: _train(self, False)     # dropout held off in both runs: .train() keeps eval mode


def _init_eval(self, *a, **k):                                    # ... and modules are born in eval mode
    _init(self, *a, **k)
    self.training = False


nn.Module.__init__ = _init_eval
with contextlib.redirect_stdout(io.StringIO()):
    import run_gun
import models
assert models.__file__.startswith(REF if which == 'reference' else pkg), models.__file__
run_gun.SummaryWriter = lambda *a, **k: types.SimpleNamespace(add_scalar=lambda *a, **k: None)
from dlsg import synth
if which == 'ours':
    from dlsg import ops, linalg as la
    from cpu_emul import CpuEmulBackend
    ops.set_backend(CpuEmulBackend())
    la.set_precision('fp32')
B, V = 2, 37
args = synth.small_args(visual_hidden_size=1024, region_projected_size=1024, query_hidden_size=1024, max_words=26, max_frames=4,
                        dataset='msr-vtt', num_obj=36, train_batch_size=B)
for k, v in dict(local_rank=0, learning_rate=1.6e-4, epoch_num=1, test_batch_size=B, save_per_epoch=1, ss_factor=20,
                 use_psl_loss=False, num_D_visual=2, lambda_D_visual=0.01).items():
    setattr(args, k, v)
train_loader, test_loader = [], []
for b in range(2):
    fr, rg, caps, lens = synth.make_inputs(B, args, V, seed=60 + b)
    train_loader.append((fr, rg, 0, caps, 0, lens, [10 * b + i for i in range(B)]))
    test_loader.append((fr, rg, 0, [100 + 10 * b + i for i in range(B)]))
torch.manual_seed(12)
random.seed(12)
np.random.seed(12)
log = io.StringIO()
with contextlib.redirect_stdout(log):
    trainer = run_gun.RunGAN(args, synth.Vocab(V), torch.device('cpu'), train_loader=train_loader, test_loader=test_loader,
                             test_reference=None, is_debug=True)
    w0 = trainer.model.decoder.word_restore.weight.detach().clone()
    f0 = trainer.D_visual.fusion.detach().clone()
    try:
        trainer.train()
    except AttributeError as e:                                   # the reference's own end_round() prints an attribute it never sets
        assert 'bleu_best' in str(e), e
sys.stdout = sys.__stdout__
np.savez(out, init_g=w0.numpy(), init_d=f0.numpy(),
         **{'g.' + k: v.detach().numpy() for k, v in trainer.model.state_dict().items()},
         **{'d.' + k: v.detach().numpy() for k, v in trainer.D_visual.state_dict().items()})
'''


BASELINE_WORKER = TRAIN_WORKER.replace("'matplotlib.pyplot', 'evaluate')", "'matplotlib.pyplot', 'evaluate', 'utils.data')")
BASELINE_WORKER = BASELINE_WORKER[:BASELINE_WORKER.index('B, V = 2, 37')].replace('import run_gun', 'import run_graph').replace(
    "run_gun.SummaryWriter = lambda *a, **k: types.SimpleNamespace(add_scalar=lambda *a, **k: None)\n", '').replace(
    "ev = sys.modules['evaluate']\n",
    "sys.modules['utils.data'].get_train_loader = sys.modules['utils.data'].get_eval_loader = None\nev = sys.modules['evaluate']\n") + r'''
B, V = 2, 37
args = synth.small_args(max_words=26, dataset='msr-vtt', train_batch_size=B)
for k, v in dict(local_rank=0, learning_rate=1.6e-4, epoch_num=2, test_batch_size=B, save_per_epoch=1, ss_factor=20).items():
    setattr(args, k, v)
train_loader, test_loader = [], []
for b in range(10):
    fr, rg, caps, lens = synth.make_inputs(B, args, V, seed=80 + b)
    train_loader.append((fr, rg, 0, caps, 0, lens, [10 * b + i for i in range(B)]))
fr, rg, _, _ = synth.make_inputs(B, args, V, seed=99)
test_loader.append((fr, rg, 0, [500, 501]))
torch.manual_seed(12)
random.seed(12)
np.random.seed(12)
log = io.StringIO()
with contextlib.redirect_stdout(log):
    trainer = run_graph.Run(args, synth.Vocab(V), torch.device('cpu'), train_loader=train_loader, test_loader=test_loader,
                            test_reference=None, is_debug=True)
    assert trainer.model.decoder.decode_hidden_size == 1300
    try:
        trainer.train()
    except AttributeError as e:
        assert 'bleu_best' in str(e), e
sys.stdout = sys.__stdout__
text = log.getvalue()
assert 'WE: ' in text and 'GT: ' in text, text[-2000:]              # the every-10-steps sample print ran (decode_tokens)
samples = [l for l in text.splitlines() if l.startswith('WE: ') or l.startswith('GT: ')]
open(out + '.txt', 'w').write('\n'.join(samples))
np.savez(out, **{'g.' + k: v.detach().numpy() for k, v in trainer.model.state_dict().items()})
'''


EDGE_WORKER = r'''
import os, sys, types, contextlib, io
ROOT, REF, which, out_path = sys.argv[1:5]
pkg = os.path.join(ROOT, 'd-lsg-video-caption_b200')
for name in ('allennlp', 'allennlp.common', 'allennlp.common.checks'):
    sys.modules[name] = types.ModuleType(name)
sys.modules['allennlp.common.checks'].ConfigurationError = type('ConfigurationError', (Exception,), {})
sys.path[:0] = [REF, pkg, ROOT] if which == 'reference' else [pkg, ROOT, os.path.join(ROOT, 'tests'), REF]
import numpy as np
import torch
torch.set_num_threads(4)
import models
assert models.__file__.startswith(REF if which == 'reference' else pkg), models.__file__
from dlsg import synth
if which == 'ours':
    from dlsg import ops, linalg as la
    from cpu_emul import CpuEmulBackend
    ops.set_backend(CpuEmulBackend())
    la.set_precision('fp32')
from models.model import CapGnnModel
res = {}
for case, args, B in (('b1', synth.small_args(), 1), ('r4', synth.small_args(num_obj=4), 3)):
    V = 37
    with contextlib.redirect_stdout(io.StringIO()):
        net = CapGnnModel(args, synth.Vocab(V))
    synth.fill_state_dict(net)
    net.eval()
    fr, rg, caps, lens = synth.make_inputs(B, args, V, seed=33)
    out, obj, mot, alpha = net(fr, rg, caps, args.max_words, 1.0)
    o = torch.cat([out[j][:lens[j]] for j in range(B)], 0)
    t = torch.cat([caps[j][:lens[j]] for j in range(B)], 0)
    torch.nn.CrossEntropyLoss()(o, t).backward()
    res[case + '.logits'], res[case + '.obj'], res[case + '.mot'] = out.detach().numpy(), obj.detach().numpy(), mot.detach().numpy()
    res[case + '.alpha'] = alpha.detach().numpy()
    for k, p in net.named_parameters():
        res[case + '.g.' + k] = np.zeros(0, np.float32) if p.grad is None else p.grad.numpy()
    with torch.no_grad():
        net.update_beam_size(3)
        res[case + '.beam3'] = net(fr, rg, None)[0].numpy()
# CapBaselineModel (model.py:76-91): the graph encoders in baseline mode (no latent pooling) + the single-head decoder over
# the 26 motion frame vectors; constructed by no live trainer, same blocks in another wiring
from models.model import CapBaselineModel
args, B, V = synth.small_args(decode_hidden_size=52), 2, 37
with contextlib.redirect_stdout(io.StringIO()):
    net = CapBaselineModel(args, synth.Vocab(V))
synth.fill_state_dict(net)
net.eval()
fr, rg, caps, lens = synth.make_inputs(B, args, V, seed=34)
logits = net(fr, rg, caps, args.max_words, 1.0)[0]
o = torch.cat([logits[j][:lens[j]] for j in range(B)], 0)
t = torch.cat([caps[j][:lens[j]] for j in range(B)], 0)
torch.nn.CrossEntropyLoss()(o, t).backward()
res['cbm.logits'] = logits.detach().numpy()
for k, p in net.named_parameters():
    res['cbm.g.' + k] = np.zeros(0, np.float32) if p.grad is None else p.grad.numpy()
with torch.no_grad():
    net.update_beam_size(1)
    res['cbm.greedy'] = net(fr, rg, None)[0].numpy()
np.savez(out_path, **res)
'''


def _run(which, out, worker=WORKER):
    r = subprocess.run([sys.executable, '-c', worker, ROOT, REF, which, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return np.load(out) if out.endswith('.npz') else __import__('json').load(open(out))


def test_reference_gather_results_gives_the_same_captions_on_our_modules(tmp_path):
    ref = _run('reference', str(tmp_path / 'ref.json'), EVAL_WORKER)
    ours = _run('ours', str(tmp_path / 'ours.json'), EVAL_WORKER)
    assert set(ref) == {'beam1', 'beam3'} and len(ref['beam1']) == 6
    assert any(len(c.split()) > 0 for c in ref['beam1'].values())
    assert ours == ref


def test_reference_train_disc_gives_the_same_critic_on_our_modules(tmp_path):
    ref = _run('reference', str(tmp_path / 'ref.npz'))
    ours = _run('ours', str(tmp_path / 'ours.npz'))
    assert abs(float(ref['loss_sum']) - float(ours['loss_sum'])) < 1e-4 * max(1.0, abs(float(ref['loss_sum'])))
    assert abs(float(ref['wass']) - float(ours['wass'])) < 1e-4 * max(1.0, abs(float(ref['wass'])))
    keys = [k for k in ref.files if k.startswith('p.')]
    assert keys == [k for k in ours.files if k.startswith('p.')]
    sys.path.insert(0, os.path.join(ROOT, 'd-lsg-video-caption_b200'))
    from dlsg import synth
    init = synth.fill_tensor('D.fusion', ref['p.fusion'].shape).numpy()
    assert np.abs(ref['p.fusion'] - init).max() > 1e-4             # the loop really trained the critic
    for k in keys:
        d = np.abs(ref[k] - ours[k])
        # three Adam steps of lr 1.6e-4: elements whose gradient sits at Adam's epsilon can differ by a fraction of a step
        assert d.max() < 2e-4, k
        assert (d > 2e-6).mean() < 2e-2, k


def test_reference_rungan_train_runs_unchanged_and_trains_the_same_weights(tmp_path):
    (tmp_path / 'ref').mkdir()
    (tmp_path / 'ours').mkdir()
    ref = _run('reference', str(tmp_path / 'ref' / 'w.npz'), TRAIN_WORKER)
    ours = _run('ours', str(tmp_path / 'ours' / 'w.npz'), TRAIN_WORKER)
    assert ref.files == ours.files and len(ref.files) > 120
    assert np.array_equal(ref['init_g'], ours['init_g']) and np.array_equal(ref['init_d'], ours['init_d'])   # same seeded init
    assert np.abs(ref['g.decoder.word_restore.weight'] - ref['init_g']).max() > 1e-4                          # both were trained
    assert np.abs(ref['d.fusion'] - ref['init_d']).max() > 1e-4
    for k in ref.files:
        d = np.abs(ref[k] - ours[k])
        # 2 generator / 4 critic Adam steps of lr 1.6e-4; elements whose gradient sits at Adam's epsilon may differ by a step
        assert d.max() < 5e-4, (k, float(d.max()))
        assert (d > 5e-6).mean() < 3e-2, (k, float((d > 5e-6).mean()))


def test_reference_baseline_trainer_runs_unchanged_and_trains_the_same_weights(tmp_path):
    (tmp_path / 'ref').mkdir()
    (tmp_path / 'ours').mkdir()
    ref = _run('reference', str(tmp_path / 'ref' / 'w.npz'), BASELINE_WORKER)
    ours = _run('ours', str(tmp_path / 'ours' / 'w.npz'), BASELINE_WORKER)
    assert ref.files == ours.files and len(ref.files) > 30
    for k in ref.files:
        d = np.abs(ref[k] - ours[k])
        # 20 Adam steps (lr 1.6e-4, halved after the first epoch)
        assert d.max() < 2e-3, (k, float(d.max()))
        assert (d > 2e-5).mean() < 3e-2, (k, float((d > 2e-5).mean()))
    # the trainer's own sample prints (arg-max tokens of the first clip, decoded by decode_tokens) are the same text
    assert open(str(tmp_path / 'ref' / 'w.npz') + '.txt').read() == open(str(tmp_path / 'ours' / 'w.npz') + '.txt').read()


def test_edge_shapes_equal_the_reference_modules(tmp_path):
    ref = _run('reference', str(tmp_path / 'ref.npz'), EDGE_WORKER)
    ours = _run('ours', str(tmp_path / 'ours.npz'), EDGE_WORKER)
    assert ref.files == ours.files
    assert not any('obj_embed' in k for k in ref.files if k.startswith('r4.'))
    for k in ref.files:
        a, b = ref[k], ours[k]
        assert a.shape == b.shape, k
        if k.endswith('.beam3') or k.endswith('.greedy'):
            assert np.array_equal(a, b), k
        elif a.size:
            assert np.abs(a - b).max() < 1e-4 * max(1.0, float(np.abs(a).max())), (k, float(np.abs(a - b).max()))
