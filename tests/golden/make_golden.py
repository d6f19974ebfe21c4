"""Generate the committed golden fixtures by running the REAL reference.

Runs only in the build container (needs /root/reference); the fixtures it writes
travel with the repo, this script is committed so they can be regenerated:

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py sites      # one family: sites|scores|topk|transfer|shipped

The reference is imported UNMODIFIED.  Its optional imports that are absent from
this image and never touched on the scoring path are stubbed as empty modules
(skimage, matplotlib, thop); its un-vendored DCT dependency `torch_dct` is
provided by oracle/torch_dct_port.py (see that file's header: parity unpinned for
the DCT arithmetic).  `load_data` (would download CIFAR) is replaced by a seeded
synthetic loader and `u2netp_inference` (unconditional `.cuda()`) by a CPU clone.

Fixture families
  sites_<net>.json        hook sessions recorded from `imp_score`: module path, hook
                          variant, hooked tensor shape, files written (+ vector length)
  scores_<tag>.npz        `imp_score` outputs on seeded inputs and seeded random-init nets
  topk_<tag>.json         (file, C, k, select_index) captured from the reference loaders'
                          own `np.argsort` calls for the README compress rates
  shipped_googlenet.npz   the 41 score files the reference ships (tie-heavy top-k inputs)
  shipped_*.npy           three of them byte-for-byte (file-format check)
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)


def import_reference():
    for name in ('skimage', 'skimage.io', 'skimage.transform', 'skimage.color',
                 'matplotlib', 'matplotlib.pyplot', 'thop'):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    sys.modules['skimage'].io = sys.modules['skimage.io']
    sys.modules['skimage'].transform = sys.modules['skimage.transform']
    sys.modules['skimage'].color = sys.modules['skimage.color']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['thop'].profile = lambda *a, **k: (0, 0)
    from oracle import torch_dct_port
    sys.modules['torch_dct'] = torch_dct_port
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import utils.common as common          # noqa: the reference, unmodified
    import utils.load_models as load_models
    return common, load_models


NET_SIDE = {'vgg_16_bn': 32, 'resnet_56': 32, 'resnet_110': 32, 'densenet_40': 32,
            'googlenet': 32, 'resnet_50': 224, 'u2netp': 320}
DATASET = {'resnet_50': 'imagenet', 'u2netp': 'DUTS'}


def synthetic_batches(batch, side, limit, u2net=False):
    """Seeded stand-in for the reference's shuffled loaders: batch b = randn(seed 1000+b)."""
    out = []
    for b in range(limit):
        g = torch.Generator().manual_seed(1000 + b)
        x = torch.randn(batch, 3, side, side, generator=g)
        out.append({'image': x} if u2net else (x, torch.zeros(batch, dtype=torch.long)))
    return out


def make_args(net, limit, batch):
    return types.SimpleNamespace(net=net, limit=limit, batch_size=batch,
                                 dataset=DATASET.get(net, 'cifar10'), data_dir='./data')


def build_net(common, net, rate=None):
    torch.manual_seed(0)
    args = types.SimpleNamespace(net=net)
    model = common.get_network(args) if rate is None else common.get_network(args, rate)
    return model.eval()


def cpu_u2netp_inference(net, loader, limit):
    net.eval()
    with torch.no_grad():
        for batch_idx, data in enumerate(loader):
            if batch_idx >= limit:
                break
            net(data['image'].type(torch.FloatTensor))


def run_imp_score(common, net_name, model, batch, side, limit):
    """Run the reference's imp_score in a temp cwd; return {file_stem: array}."""
    loader = synthetic_batches(batch, side, limit, u2net=(net_name == 'u2netp'))
    common.load_data = lambda args: (loader, None)
    common.u2netp_inference = cpu_u2netp_inference
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            common.imp_score(model, make_args(net_name, limit, batch))
            d = os.path.join(tmp, 'importance_score', '%s_limit%d' % (net_name, limit))
            out = {}
            for fn in sorted(os.listdir(d)):
                with open(os.path.join(d, fn), 'rb') as f:
                    raw = f.read()
                arr = np.load(os.path.join(d, fn))
                assert arr.dtype == np.float32 and arr.ndim == 1
                assert len(raw) == 128 + 4 * arr.shape[0], (fn, len(raw))
                out[fn[:-4]] = arr
        finally:
            os.chdir(cwd)
    return out


# ------------------------------------------------------------------------- sites
def gen_sites(common):
    for net_name, side in NET_SIDE.items():
        model = build_net(common, net_name)
        paths = {id(m): n for n, m in model.named_modules()}
        sessions = []

        def recorder(variant):
            def hook(module, inputs, output):
                t = inputs[0] if variant == 'I' else output
                sessions.append({'module': paths[id(module)], 'variant': variant,
                                 'shape': list(t.shape[1:]), 'files': []})
            return hook

        saved = (common.get_feature_hook, common.get_feature_hook_densenet,
                 common.get_feature_hook_u2net_input, common.np.save)
        common.get_feature_hook = recorder('O')
        common.get_feature_hook_densenet = recorder('D')
        common.get_feature_hook_u2net_input = recorder('I')

        class NpProxy:
            def __getattr__(self, k):
                return getattr(np, k)

            @staticmethod
            def save(path, arr):
                sessions[-1]['files'].append(os.path.basename(path)[:-4])
        real_np = common.np
        common.np = NpProxy()
        loader = synthetic_batches(1, side, 1, u2net=(net_name == 'u2netp'))
        common.load_data = lambda args: (loader, None)
        common.u2netp_inference = cpu_u2netp_inference
        # the recorder leaves feature_result a 0-dim tensor; slicing it (googlenet) needs a vector
        orig_tensor = common.torch.tensor
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                if net_name == 'googlenet':
                    common.torch = types.SimpleNamespace(
                        **{k: getattr(torch, k) for k in dir(torch) if not k.startswith('__')})
                    common.torch.tensor = lambda *a, **k: torch.zeros(4096)
                common.imp_score(model, make_args(net_name, 1, 1))
            finally:
                os.chdir(cwd)
                common.torch = torch
                common.np = real_np
                (common.get_feature_hook, common.get_feature_hook_densenet,
                 common.get_feature_hook_u2net_input, _) = saved
        with open(os.path.join(HERE, 'sites_%s.json' % net_name), 'w') as f:
            json.dump({'net': net_name, 'input_side': side, 'sessions': sessions}, f, indent=0)
        print('sites', net_name, len(sessions), 'sessions',
              sum(len(s['files']) for s in sessions), 'files')


# ------------------------------------------------------------------------ scores
SCORE_CASES = [  # tag, net, batch, side, limit
    ('vgg_16_bn_b3_l2', 'vgg_16_bn', 3, 32, 2),
    ('resnet_56_b2_l2', 'resnet_56', 2, 32, 2),
    ('resnet_110_b1_l1', 'resnet_110', 1, 32, 1),
    ('densenet_40_b2_l1', 'densenet_40', 2, 32, 1),
    ('googlenet_b2_l1', 'googlenet', 2, 32, 1),
    ('resnet_50_s64_b2_l1', 'resnet_50', 2, 64, 1),
    ('resnet_50_s224_b1_l1', 'resnet_50', 1, 224, 1),
    ('u2netp_s64_b1_l2', 'u2netp', 1, 64, 2),
    ('u2netp_s144_b1_l1', 'u2netp', 1, 144, 1),
    # the headline sizes themselves (round 2): BASELINE config 5's 320x320, the 288x288 the reference's DUTS loader really
    # feeds (utils/common.py:154-155), and ResNet-50@224 with more than one image
    ('u2netp_s288_b1_l1', 'u2netp', 1, 288, 1),
    ('u2netp_s320_b1_l1', 'u2netp', 1, 320, 1),
    ('resnet_50_s224_b2_l1', 'resnet_50', 2, 224, 1),
]


def state_digest(model):
    h = hashlib.sha256()
    for k, v in model.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def gen_scores(common, only=None):
    for tag, net_name, batch, side, limit in SCORE_CASES:
        if only and only not in tag:
            continue
        model = build_net(common, net_name)
        digest = state_digest(model)
        out = run_imp_score(common, net_name, model, batch, side, limit)
        meta = json.dumps({'net': net_name, 'batch': batch, 'side': side, 'limit': limit,
                           'weights_sha256': digest, 'seed': 0, 'batch_seed_base': 1000})
        np.savez_compressed(os.path.join(HERE, 'scores_%s.npz' % tag), __meta__=np.array(meta), **out)
        print('scores', tag, len(out), 'files')


# ------------------------------------------------------------------------- top-k
TOPK_CASES = [  # tag, net, README compress rate (README.md:90,114,138,162,186,211; prune_u2netp.py:99)
    ('vgg_16_bn', 'vgg_16_bn_b3_l2', '[0.50]*7+[0.95]*5'),
    ('resnet_56', 'resnet_56_b2_l2', '[0.]+[0.18]*29'),
    ('resnet_110', 'resnet_110_b1_l1', '[0.]+[0.2]*2+[0.3]*18+[0.40]*18+[0.39]*19'),
    ('densenet_40', 'densenet_40_b2_l1', '[0.]+[0.2]*12+[0.]+[0.2]*12+[0.]+[0.2]*12'),
    ('googlenet', 'googlenet_b2_l1', '[0.4]+[0.85]*2+[0.9]*5+[0.9]*2'),
    ('resnet_50', 'resnet_50_s64_b2_l1', '[0.]+[0.1]*3+[0.4]*7+[0.4]*9'),
    ('u2netp', 'u2netp_s64_b1_l2', '[0.40]*40'),
]


def gen_topk(common, load_models):
    for net_name, score_tag, rate_str in TOPK_CASES:
        rate = common.get_compress_rate(types.SimpleNamespace(compress_rate=rate_str))
        orig = build_net(common, net_name)
        pruned = build_net(common, net_name, rate)
        scores = np.load(os.path.join(HERE, 'scores_%s.npz' % score_tag))
        calls = []
        with tempfile.TemporaryDirectory() as tmp:
            for k in scores.files:
                if k != '__meta__':
                    np.save(os.path.join(tmp, k + '.npy'), scores[k])

            class NpProxy:
                def __getattr__(self, k):
                    return getattr(np, k)

                @staticmethod
                def load(path):
                    arr = np.load(path)
                    calls.append({'file': os.path.basename(path)[:-4], 'C': int(arr.shape[0])})
                    return arr

                @staticmethod
                def argsort(a, *args, **kw):
                    res = np.argsort(a, *args, **kw)

                    class Tap(np.ndarray):
                        def __getitem__(self, item):
                            got = np.asarray(self).__getitem__(item)
                            if isinstance(item, slice) and 'k' not in calls[-1]:
                                calls[-1]['k'] = int(len(got))
                                calls[-1]['select_index'] = sorted(int(v) for v in got)
                            return got
                    return res.view(Tap)
            load_models.np = NpProxy()
            args = types.SimpleNamespace(imp_score=tmp, net=net_name)
            try:
                od = orig.state_dict()
                if net_name == 'vgg_16_bn':
                    load_models.load_vgg_model(pruned, od, args)
                elif net_name == 'resnet_56':
                    load_models.load_resnet_model(pruned, od, 56, args)
                elif net_name == 'resnet_110':
                    load_models.load_resnet_model(pruned, od, 110, args)
                elif net_name == 'densenet_40':
                    load_models.load_densenet_model(pruned, od, args)
                elif net_name == 'googlenet':
                    load_models.load_google_model(pruned, od, args)
                elif net_name == 'resnet_50':
                    load_models.load_resnet_imagenet_model(pruned, od, args)
                elif net_name == 'u2netp':
                    load_models.load_u2netp_model(pruned, od, args)
            finally:
                load_models.np = np
        with open(os.path.join(HERE, 'topk_%s.json' % net_name), 'w') as f:
            json.dump({'net': net_name, 'scores': score_tag, 'compress_rate': rate_str,
                       'rates': rate, 'selections': calls}, f)
        print('topk', net_name, len(calls), 'selections')


# ---------------------------------------------------------------------- transfer
def tensor_digests(state_dict):
    out = {}
    for k, v in state_dict.items():
        a = v.detach().cpu().contiguous().numpy()
        out[k] = [list(a.shape), str(a.dtype), hashlib.sha256(a.tobytes()).hexdigest()]
    return out


def gen_transfer(common, load_models, only=None):
    """Pruned-weight transfer: run the reference loaders unmodified and record a digest of every tensor of
    the pruned model they produce (orig net: seed 0, rate 0; pruned net: seed 0, README rate; scores: scores_<tag>.npz)."""
    for net_name, score_tag, rate_str in TOPK_CASES:
        if only and only != net_name:
            continue
        rate = common.get_compress_rate(types.SimpleNamespace(compress_rate=rate_str))
        orig = build_net(common, net_name)
        pruned = build_net(common, net_name, rate)
        before = tensor_digests(pruned.state_dict())
        scores = np.load(os.path.join(HERE, 'scores_%s.npz' % score_tag))
        with tempfile.TemporaryDirectory() as tmp:
            for k in scores.files:
                if k != '__meta__':
                    np.save(os.path.join(tmp, k + '.npy'), scores[k])
            args = types.SimpleNamespace(imp_score=tmp, net=net_name)
            od = orig.state_dict()
            if net_name == 'vgg_16_bn':
                load_models.load_vgg_model(pruned, od, args)
            elif net_name == 'resnet_56':
                load_models.load_resnet_model(pruned, od, 56, args)
            elif net_name == 'resnet_110':
                load_models.load_resnet_model(pruned, od, 110, args)
            elif net_name == 'densenet_40':
                load_models.load_densenet_model(pruned, od, args)
            elif net_name == 'googlenet':
                load_models.load_google_model(pruned, od, args)
            elif net_name == 'resnet_50':
                load_models.load_resnet_imagenet_model(pruned, od, args)
            elif net_name == 'u2netp':
                load_models.load_u2netp_model(pruned, od, args)
        after = tensor_digests(pruned.state_dict())
        # kept small: one digest of the pruned model before the loader, per-tensor digests only where the loader wrote
        h = hashlib.sha256()
        for k in before:
            h.update(k.encode())
            h.update(before[k][2].encode())
        changed = {k: after[k] for k in after if after[k][2] != before[k][2]}
        same = hashlib.sha256(''.join(k + after[k][2] for k in after if after[k][2] == before[k][2]).encode()).hexdigest()
        with open(os.path.join(HERE, 'transfer_%s.json' % net_name), 'w') as f:
            json.dump({'net': net_name, 'scores': score_tag, 'compress_rate': rate_str, 'pruned_init_digest': h.hexdigest(),
                       'n_tensors': len(after), 'changed': changed, 'unchanged_digest': same}, f)
        print('transfer', net_name, len(after), 'tensors,', len(changed), 'changed by the loader')


def gen_transfer_iter(common, load_models):
    """One round of prune_dynamic.py:150-154 for GoogLeNet: score the net pruned at 1/5 of the README rate, then let the
    reference loader (with its `cpr` argument = the rates of that already pruned net) fill the net pruned at 2/5."""
    final = common.get_compress_rate(types.SimpleNamespace(compress_rate='[0.4]+[0.85]*2+[0.9]*5+[0.9]*2'))
    rate_a, rate_b = list(np.array(final) / 5), list((np.array(final) / 5) * 2)
    net_a = build_net(common, 'googlenet', rate_a)
    net_b = build_net(common, 'googlenet', rate_b)
    before = tensor_digests(net_b.state_dict())
    scores = run_imp_score(common, 'googlenet', net_a, 2, 32, 1)
    calls = []
    with tempfile.TemporaryDirectory() as tmp:
        for k, v in scores.items():
            np.save(os.path.join(tmp, k + '.npy'), v)

        class NpProxy:
            def __getattr__(self, k):
                return getattr(np, k)

            @staticmethod
            def load(path):
                arr = np.load(path)
                calls.append({'file': os.path.basename(path)[:-4], 'C': int(arr.shape[0])})
                return arr

            @staticmethod
            def argsort(a, *args, **kw):
                res = np.argsort(a, *args, **kw)

                class Tap(np.ndarray):
                    def __getitem__(self, item):
                        got = np.asarray(self).__getitem__(item)
                        if isinstance(item, slice) and 'k' not in calls[-1]:
                            calls[-1]['k'] = int(len(got))
                            calls[-1]['select_index'] = sorted(int(v) for v in got)
                        return got
                return res.view(Tap)
        load_models.np = NpProxy()
        try:
            load_models.load_google_model(net_b, net_a.state_dict(), types.SimpleNamespace(imp_score=tmp, net='googlenet'), rate_a)
        finally:
            load_models.np = np
    after = tensor_digests(net_b.state_dict())
    h = hashlib.sha256()
    for k in before:
        h.update(k.encode())
        h.update(before[k][2].encode())
    changed = {k: after[k] for k in after if after[k][2] != before[k][2]}
    same = hashlib.sha256(''.join(k + after[k][2] for k in after if after[k][2] == before[k][2]).encode()).hexdigest()
    np.savez_compressed(os.path.join(HERE, 'scores_googlenet_iter_b2_l1.npz'), **scores)
    with open(os.path.join(HERE, 'transfer_googlenet_iter.json'), 'w') as f:
        json.dump({'net': 'googlenet', 'origin_rates': rate_a, 'rates': rate_b, 'scores': 'googlenet_iter_b2_l1',
                   'pruned_init_digest': h.hexdigest(), 'n_tensors': len(after), 'changed': changed, 'unchanged_digest': same,
                   'selections': calls}, f)
    print('transfer_iter googlenet', len(after), 'tensors,', len(changed), 'changed,', len(calls), 'selections')


TRANSFER_ALT_CASES = [  # a second compress rate per net (README.md's other settings): different k patterns, same scores
    ('vgg_16_bn', 'vgg_16_bn_b3_l2', '[0.30]*7+[0.75]*5'),
    ('resnet_56', 'resnet_56_b2_l2', '[0.]+[0.4]*2+[0.5]*9+[0.6]*9+[0.7]*9'),
    ('resnet_110', 'resnet_110_b1_l1', '[0.]+[0.4]*2+[0.5]*18+[0.65]*36'),
    ('densenet_40', 'densenet_40_b2_l1', '[0.]+[0.4]*12+[0.]+[0.4]*12+[0.]+[0.4]*12'),
    ('googlenet', 'googlenet_b2_l1', '[0.3]+[0.6]*2+[0.7]*5+[0.8]*2'),
    ('resnet_50', 'resnet_50_s64_b2_l1', '[0.]+[0.2]*3+[0.65]*16'),
    ('u2netp', 'u2netp_s64_b1_l2', '[0.20]*40'),
]


def gen_transfer_alt(common, load_models, only=None):
    loaders = {'vgg_16_bn': lambda m, od, a: load_models.load_vgg_model(m, od, a),
               'resnet_56': lambda m, od, a: load_models.load_resnet_model(m, od, 56, a),
               'resnet_110': lambda m, od, a: load_models.load_resnet_model(m, od, 110, a),
               'densenet_40': lambda m, od, a: load_models.load_densenet_model(m, od, a),
               'googlenet': lambda m, od, a: load_models.load_google_model(m, od, a),
               'resnet_50': lambda m, od, a: load_models.load_resnet_imagenet_model(m, od, a),
               'u2netp': lambda m, od, a: load_models.load_u2netp_model(m, od, a)}
    for net_name, score_tag, rate_str in TRANSFER_ALT_CASES:
        if only and only != net_name:
            continue
        rate = common.get_compress_rate(types.SimpleNamespace(compress_rate=rate_str))
        orig = build_net(common, net_name)
        pruned = build_net(common, net_name, rate)
        before = tensor_digests(pruned.state_dict())
        scores = np.load(os.path.join(HERE, 'scores_%s.npz' % score_tag))
        calls = []
        with tempfile.TemporaryDirectory() as tmp:
            for k in scores.files:
                if k != '__meta__':
                    np.save(os.path.join(tmp, k + '.npy'), scores[k])

            class NpProxy:
                def __getattr__(self, k):
                    return getattr(np, k)

                @staticmethod
                def load(path):
                    arr = np.load(path)
                    calls.append({'file': os.path.basename(path)[:-4], 'C': int(arr.shape[0])})
                    return arr

                @staticmethod
                def argsort(a, *args, **kw):
                    res = np.argsort(a, *args, **kw)

                    class Tap(np.ndarray):
                        def __getitem__(self, item):
                            got = np.asarray(self).__getitem__(item)
                            if isinstance(item, slice) and 'k' not in calls[-1]:
                                calls[-1]['k'] = int(len(got))
                                calls[-1]['select_index'] = sorted(int(v) for v in got)
                            return got
                    return res.view(Tap)
            load_models.np = NpProxy()
            try:
                loaders[net_name](pruned, orig.state_dict(), types.SimpleNamespace(imp_score=tmp, net=net_name))
            finally:
                load_models.np = np
        after = tensor_digests(pruned.state_dict())
        h = hashlib.sha256()
        for k in before:
            h.update(k.encode())
            h.update(before[k][2].encode())
        changed = {k: after[k] for k in after if after[k][2] != before[k][2]}
        same = hashlib.sha256(''.join(k + after[k][2] for k in after if after[k][2] == before[k][2]).encode()).hexdigest()
        with open(os.path.join(HERE, 'transfer_%s_alt.json' % net_name), 'w') as f:
            json.dump({'net': net_name, 'scores': score_tag, 'compress_rate': rate_str, 'pruned_init_digest': h.hexdigest(),
                       'n_tensors': len(after), 'changed': changed, 'unchanged_digest': same, 'selections': calls}, f)
        print('transfer_alt', net_name, len(after), 'tensors,', len(changed), 'changed,', len(calls), 'selections')


# ----------------------------------------------------------------------- shipped
def gen_shipped():
    out = {}
    for sub in ('googlenet_limit5', 'googlenet_limit1'):
        d = os.path.join(REF, 'importance_score', sub)
        for fn in sorted(os.listdir(d)):
            out['%s/%s' % (sub, fn[:-4])] = np.load(os.path.join(d, fn))
    np.savez_compressed(os.path.join(HERE, 'shipped_googlenet.npz'), **out)
    for sub, fn in (('googlenet_limit5', 'imp_conv2_n3x3.npy'), ('googlenet_limit5', 'imp_conv1_.npy'),
                    ('googlenet_limit1', 'imp_conv2_n5x5.npy')):
        with open(os.path.join(REF, 'importance_score', sub, fn), 'rb') as f:
            raw = f.read()
        with open(os.path.join(HERE, 'shipped_%s_%s' % (sub, fn)), 'wb') as f:
            f.write(raw)
    print('shipped', len(out), 'arrays')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    only = sys.argv[2] if len(sys.argv) > 2 else None
    torch.set_num_threads(8)
    common, load_models = import_reference()
    if what in ('all', 'sites'):
        gen_sites(common)
    if what in ('all', 'scores'):
        gen_scores(common, only)
    if what in ('all', 'topk'):
        gen_topk(common, load_models)
    if what in ('all', 'transfer'):
        gen_transfer(common, load_models, only)
    if what in ('all', 'transfer_iter'):
        gen_transfer_iter(common, load_models)
    if what in ('all', 'transfer_alt'):
        gen_transfer_alt(common, load_models, only)
    if what in ('all', 'shipped'):
        gen_shipped()
