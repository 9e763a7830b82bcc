"""Import the UNMODIFIED reference (MrZihan/GridMM) in-process, CPU only.

TEST INFRASTRUCTURE -- only `oracle/make_golden.py` and tests that are skipped
when `/root/reference` is absent may use this.  Nothing on the product path, in
`-m gpu` tests, `smoke()` or `bench.py` imports it: `/root/reference` does not
exist on the GPU box.

The reference needs four harness-side shims and zero source edits:
  1. stub modules for imports that are absent here and unused by the hot path
     (`easydict`, `MatterSim`, `h5py`, `imutils`, `jsonlines`, `line_profiler`,
     `cv2`/`PIL` are real);
  2. `GlocalTextPathNavCMT.init_weights` replaced by a no-op, because
     transformers 5.x expects `post_init()` (map_nav_src/models/vilmodel.py:712);
     the config is built from `transformers.BertConfig()` + the attributes that
     map_nav_src/models/vlnbert_init.py:38-56 sets;
  3. `EnvBatch` is created with `object.__new__` and fed fake simulator state,
     depth DB, semantic DB and viewpoint_info (all that
     map_nav_src/r2r/env.py:267-374 touches);
  4. after each `getGlobalMap` the accumulated feature array is re-viewed as an
     ndarray subclass whose `== []` is False, because
     `if self.global_semantic[i] == []` (env.py:298) raises under numpy >= 2.
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("GRIDMM_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "map_nav_src")


def available():
    return os.path.isdir(REF_SRC)


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_stubs():
    try:
        import easydict  # noqa: F401
    except ImportError:
        _stub("easydict", EasyDict=_AttrDict)
    for name in ("MatterSim", "h5py", "imutils", "jsonlines", "line_profiler"):
        try:
            __import__(name)
        except ImportError:
            _stub(name)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)


_MODEL_CACHE = {}


def make_config(obj_feat_size=0, num_l_layers=9, num_pano_layers=2, num_x_layers=4,
                image_feat_size=768, angle_feat_size=4, glocal_fuse=True, graph_sprels=True):
    """BertConfig + the attributes of map_nav_src/models/vlnbert_init.py:38-56."""
    from transformers import BertConfig
    c = BertConfig()
    c.max_action_steps = 100
    c.image_feat_size = image_feat_size
    c.angle_feat_size = angle_feat_size
    c.obj_feat_size = obj_feat_size
    c.obj_loc_size = 3
    c.num_l_layers = num_l_layers
    c.num_pano_layers = num_pano_layers
    c.num_x_layers = num_x_layers
    c.graph_sprels = graph_sprels
    c.glocal_fuse = glocal_fuse
    c.fix_lang_embedding = False
    c.fix_pano_embedding = False
    c.fix_local_branch = False
    c.update_lang_bert = True
    c.output_attentions = True
    c.pred_head_dropout_prob = 0.1
    c.use_lang2visn_attn = False
    return c


def load_reference_model(seed=0, **cfg):
    """Build the reference GlocalTextPathNavCMT with deterministic N(0,0.02) weights."""
    import torch
    _install_stubs()
    from models import vilmodel as ref_vilmodel  # noqa: E402  (reference module)
    cls = ref_vilmodel.GlocalTextPathNavCMT
    cls.init_weights = lambda self: None
    config = make_config(**cfg)
    model = cls(config)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1 and name.endswith(".weight"):
                # LayerNorm gamma: 1 + jitter so a gamma bug cannot hide
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    model.eval()
    return model


class _NeverEqList(np.ndarray):
    """ndarray whose comparison with a list is a plain False (shim 4)."""

    def __eq__(self, other):
        if isinstance(other, list):
            return False
        return np.ndarray.__eq__(self, other)

    __hash__ = None


class _FakeLoc:
    def __init__(self, vp):
        self.viewpointId = vp


class _FakeState:
    def __init__(self, scan, vp, heading):
        self.scanId = scan
        self.location = _FakeLoc(vp)
        self.heading = heading


class _FakeSim:
    def __init__(self):
        self.state = None

    def getState(self):
        return [self.state]


class _DictDB:
    def __init__(self):
        self.store = {}

    def get_image_feature(self, scan, vp):
        return self.store["%s_%s" % (scan, vp)]


def load_reference_env(batch_size):
    """A reference EnvBatch (map_nav_src/r2r/env.py:125) without MatterSim/h5py."""
    _install_stubs()
    import importlib
    # r2r/env.py imports `utils.data` and `r2r.eval_utils`; both import cleanly with stubs
    env_mod = importlib.import_module("r2r.env")
    EnvBatch = env_mod.EnvBatch
    env = object.__new__(EnvBatch)
    env.batch_size = batch_size
    env.sims = [_FakeSim() for _ in range(batch_size)]
    env.DepthDB = _DictDB()
    env.SemanticDB = _DictDB()
    env.viewpoint_info = {}
    env.feature_states = [None] * batch_size
    # the state lists of EnvBatch.__init__/newEpisodes (env.py:142-151, 183-193)
    env.global_semantic = [[] for _ in range(batch_size)]
    env.global_position_x = [[] for _ in range(batch_size)]
    env.global_position_y = [[] for _ in range(batch_size)]
    env.global_mask = [[] for _ in range(batch_size)]
    env.max_x = [-10000 for _ in range(batch_size)]
    env.min_x = [10000 for _ in range(batch_size)]
    env.max_y = [-10000 for _ in range(batch_size)]
    env.min_y = [10000 for _ in range(batch_size)]
    env.heading = [0 for _ in range(batch_size)]
    env.global_map = [[] for _ in range(batch_size)]
    return env, env_mod


def ref_step(env, i, scan, vp, heading, depth_u16, clip_fp16, pos_xy):
    """Feed one viewpoint to the reference getGlobalMap(i) and return its outputs.

    depth_u16: uint16[36,128,128]; clip_fp16: float16[12, 50, 768]; pos_xy: python floats.
    Returns (grid_fts fp16[N,768], grid_map f64[N], gridmap_pos_fts f32[196,5]).
    """
    key = "%s_%s" % (scan, vp)
    env.DepthDB.store[key] = depth_u16
    env.SemanticDB.store[key] = clip_fp16
    env.viewpoint_info[key] = {"x": float(pos_xy[0]), "y": float(pos_xy[1]), "z": 0.0}
    env.sims[i].state = _FakeState(scan, vp, heading)
    out = env.getGlobalMap(i)
    # mirror getStates (env.py:397): write the returned state back, then shim 4
    (_, env.global_semantic[i], env.global_position_x[i], env.global_position_y[i],
     env.global_mask[i], env.global_map[i], env.max_x[i], env.min_x[i], env.max_y[i],
     env.min_y[i], gridmap_pos_fts) = out
    env.global_semantic[i] = np.asarray(env.global_semantic[i]).view(_NeverEqList)
    return np.asarray(env.global_semantic[i]), np.array(env.global_map[i]), gridmap_pos_fts


# ----------------------------------------------------------------------------------------------- pretraining dataset variant
def load_reference_pretrain_data():
    """The pretraining dataset object (pretrain_src/data/dataset.py:90 `ReverieTextPathData`, subclass `R2RTextPathData` :634)
    created with `object.__new__` and fake depth / CLIP / viewpoint stores -- all that its `getGlobalMap` (:351-473) touches.
    The module is imported under the alias package `pretrain_data` so that its `.common` import resolves and nothing in
    map_nav_src is shadowed."""
    import importlib
    _install_stubs()
    if "pretrain_data" not in sys.modules:
        pkg = types.ModuleType("pretrain_data")
        pkg.__path__ = [os.path.join(REF_ROOT, "pretrain_src", "data")]
        sys.modules["pretrain_data"] = pkg
    mod = importlib.import_module("pretrain_data.dataset")
    ds = object.__new__(mod.R2RTextPathData)
    ds.DepthDB = _DictDB()
    ds.SemanticDB = _DictDB()
    ds.viewpoint_info = {}
    return ds


def ref_pretrain_traj(ds, scan, headings, depth_u16, clip_f32, pos_xy):
    """The grid half of `get_traj_pano_fts` (pretrain_src/data/dataset.py:482-507) over one ground-truth path of T viewpoints:
    state reset, one `getGlobalMap` per viewpoint, features concatenated by the caller.  `headings[t]` is what :499 would have
    set from the candidate view index ((viewidx % 12) * 30 degrees).
    Returns per step: grid_map f64[588 (t+1)], gridmap_pos_fts f32[196,5], target_patch_id; and the final grid_fts [588 T, D]."""
    T = len(headings)
    path = ["vp%d" % t for t in range(T)]
    for t in range(T):
        key = "%s_%s" % (scan, path[t])
        ds.DepthDB.store[key] = depth_u16[t]
        ds.SemanticDB.store[key] = clip_f32[t]
        ds.viewpoint_info[key] = {"x": float(pos_xy[t][0]), "y": float(pos_xy[t][1]), "z": 0.0}
    ds.gt_path = path
    ds.global_semantic, ds.global_position_x, ds.global_position_y, ds.global_mask = [], [], [], []   # :483-491
    ds.max_x, ds.min_x, ds.max_y, ds.min_y = -10000, 10000, -10000, 10000
    ds.global_map = None
    cells, pos_fts, targets = [], [], []
    grid_fts = np.array([])
    for t in range(T):
        ds.heading = headings[t]
        out = ds.getGlobalMap(scan, path[t])
        (sem, ds.global_position_x, ds.global_position_y, ds.global_mask, ds.global_map, ds.max_x, ds.min_x, ds.max_y, ds.min_y,
         pf, target) = out
        ds.global_semantic = np.asarray(sem).view(_NeverEqList)       # shim 4 (`== []` under numpy >= 2, dataset.py:388)
        grid_fts = np.asarray(sem) if grid_fts.shape == (0,) else np.concatenate((grid_fts, np.asarray(sem)), axis=0)   # :503-506
        cells.append(np.array(ds.global_map))
        pos_fts.append(pf)
        targets.append(int(target))
    return cells, pos_fts, targets, grid_fts.reshape((-1, clip_f32[0].shape[-1]))


def make_pretrain_config(num_l_layers=9, num_pano_layers=2, num_x_layers=4, obj_feat_size=0):
    """BertConfig + the keys of pretrain_src/config/r2r_model_config.json (obj_feat_size > 0: the REVERIE / SOON configs)."""
    from transformers import BertConfig
    c = BertConfig()
    for k, v in dict(pred_head_dropout_prob=0.1, image_feat_size=768, image_prob_size=1000, angle_feat_size=4, obj_feat_size=obj_feat_size,
                     obj_prob_size=0, num_l_layers=num_l_layers, num_x_layers=num_x_layers, num_pano_layers=num_pano_layers,
                     max_action_steps=100, update_lang_bert=True, use_lang2visn_attn=True, graph_sprels=True,
                     glocal_fuse=True).items():
        setattr(c, k, v)
    return c


def load_reference_pretrain_model(**cfg):
    """The pretraining trunk `GlocalTextPathCMT` (pretrain_src/model/vilmodel.py:640-855), eval mode, imported under the alias
    package `pretrain_model`; `init_weights` is a no-op as for the navigation model (shim 2)."""
    import importlib
    _install_stubs()
    if "pretrain_model" not in sys.modules:
        pkg = types.ModuleType("pretrain_model")
        pkg.__path__ = [os.path.join(REF_ROOT, "pretrain_src", "model")]
        sys.modules["pretrain_model"] = pkg
    vil = importlib.import_module("pretrain_model.vilmodel")
    cls = vil.GlocalTextPathCMT
    cls.init_weights = lambda self: None
    model = cls(make_pretrain_config(**cfg))
    model.eval()
    return model


def load_reference_pretrain_wrapper(tasks=("mlm", "sap"), **cfg):
    """`GlocalTextPathCMTPreTraining` (pretrain_src/model/pretrain_cmt.py:37-66): the trunk as `.bert` plus the task heads.
    `init_weights` / `tie_weights` are no-ops under transformers 5.x (shim 2); the caller ties the MLM decoder to the word
    embeddings by value, which is all `tie_weights` (:68-71) achieves for a forward pass."""
    import importlib
    load_reference_pretrain_model(num_l_layers=1, num_pano_layers=1, num_x_layers=1)      # installs the alias package
    pc = importlib.import_module("pretrain_model.pretrain_cmt")
    cls = pc.GlocalTextPathCMTPreTraining
    cls.init_weights = lambda self: None
    cls.tie_weights = lambda self, *a, **k: None
    config = make_pretrain_config(**cfg)
    config.pretrain_tasks = list(tasks)
    model = cls(config)
    model.eval()
    return model


# ----------------------------------------------------------------------------------------------- continuous-env variant
CE_ROOT = os.path.join(REF_ROOT, "VLN_CE", "vlnce_baselines", "models")


def _extract_functions(path, names, class_name=None):
    """AST-extract function definitions (module level, or methods of `class_name`) from a reference file WITHOUT importing
    the module (the CE policy file imports habitat / gym / timm, which are absent here).  Returns source-compiled objects."""
    import ast
    tree = ast.parse(open(path).read(), filename=path)
    body = tree.body
    if class_name is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise RuntimeError("%s lacks %s" % (path, sorted(missing)))
    return picked


def load_reference_ce_grid(batch_size, dataset="R2R", max_dist=25):
    """The reference's continuous-env grid builder: `get_rel_position`, `get_gridmap_pos_fts`, `getGlobalMap` of
    VLN_CE/vlnce_baselines/models/Policy_ViewSelection_GridMap.py (:632-641, 661-684, 689-825) and the two helpers of
    VLN_CE/vlnce_baselines/models/utils.py (:110-144), compiled from the reference's own source text into a bare class."""
    import ast
    import math
    pol = os.path.join(CE_ROOT, "Policy_ViewSelection_GridMap.py")
    utl = os.path.join(CE_ROOT, "utils.py")
    helpers = _extract_functions(utl, ["calculate_vp_rel_pos_fts", "get_angle_fts"])
    cls_name = next(n.name for n in ast.parse(open(pol).read()).body
                    if isinstance(n, ast.ClassDef) and any(isinstance(m, ast.FunctionDef) and m.name == "getGlobalMap" for m in n.body))
    methods = _extract_functions(pol, ["get_rel_position", "get_gridmap_pos_fts", "getGlobalMap"], class_name=cls_name)
    holder = ast.ClassDef(name="RefCEGrid", bases=[], keywords=[], body=methods, decorator_list=[])
    mod = ast.Module(body=helpers + [holder], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np, "math": math, "DATASET": dataset, "MAX_DIST": max_dist}
    exec(compile(mod, pol, "exec"), ns)
    g = ns["RefCEGrid"]()
    g.global_fts = [[] for _ in range(batch_size)]
    g.global_position_x = [[] for _ in range(batch_size)]
    g.global_position_y = [[] for _ in range(batch_size)]
    g.global_mask = [[] for _ in range(batch_size)]
    g.global_map_index = [[] for _ in range(batch_size)]
    g.max_x = [-10000 for _ in range(batch_size)]
    g.min_x = [10000 for _ in range(batch_size)]
    g.max_y = [-10000 for _ in range(batch_size)]
    g.min_y = [10000 for _ in range(batch_size)]
    g.headings = [0.0 for _ in range(batch_size)]
    return g


def ref_ce_step(g, i, heading, depth_f32, clip_fp16, pos_xy):
    """One reference CE getGlobalMap call.  depth_f32: float32[12,256,256] metres; clip_fp16: [12,50,768]."""
    g.headings[i] = heading
    out = g.getGlobalMap(i, {"x": float(pos_xy[0]), "y": float(pos_xy[1])}, heading, depth_f32, clip_fp16, None)
    g.global_fts[i] = np.asarray(out[0]).view(_NeverEqList)          # shim 4 (`== []` under numpy >= 2)
    return np.asarray(out[0]), np.array(out[4]), out[9]


def load_reference_ce_model(**cfg):
    """VLN_CE/vlnce_baselines/models/gridmap/vilmodel.py GlocalTextPathNavCMT (the CE copy).  Its constructor builds an online
    CLIP (pure torch, importable) and `timm.create_model(...)` (timm is absent: stubbed with an empty module -- neither is
    touched by forward('navigation'))."""
    import importlib
    import torch
    _install_stubs()
    from transformers import BertConfig, BertPreTrainedModel  # noqa: F401  (resolve transformers' lazy imports BEFORE timm is stubbed)
    if "timm" not in sys.modules:
        t = _stub("timm", create_model=lambda *a, **k: torch.nn.Identity())
        d = _stub("timm.data", resolve_data_config=lambda *a, **k: {})
        tf = _stub("timm.data.transforms_factory", create_transform=lambda *a, **k: None)
        t.data = d
        d.transforms_factory = tf
    pkg = types.ModuleType("ce_gridmap")
    pkg.__path__ = [os.path.join(CE_ROOT, "gridmap")]
    sys.modules["ce_gridmap"] = pkg
    vil = importlib.import_module("ce_gridmap.vilmodel")
    cls = vil.GlocalTextPathNavCMT
    cls.init_weights = lambda self: None
    kw = dict(obj_feat_size=0)
    kw.update(cfg)
    model = cls(make_config(**kw))
    model.eval()
    return model
