"""patch_reference(): rebind the reference's hot-path callables to the B200 implementations so that its own
entry points (train.py, test.py, config/*.yaml, the trainer classes) run unmodified.

    import openess_b200.patch as p; p.patch_reference("/path/to/OpenESS")   # before `import training...`

or `OPENESS_B200=1 python train.py ...` with the sitecustomize stub shown in INTEGRATION.md."""
import importlib.util
import os
import sys
import types

_PATCHED = {}


def _load(ref_root, name, rel):
    """Import a reference module by file path under its reference name (the repo has namespace packages that are
    shadowed by unrelated installed packages, e.g. `datasets`)."""
    if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(ref_root):
        return sys.modules[name]
    parts = name.split(".")
    for i in range(1, len(parts)):
        pkg = ".".join(parts[:i])
        if pkg not in sys.modules or not str(getattr(sys.modules[pkg], "__path__", [""])[0]).startswith(ref_root):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(ref_root, *parts[:i])]
            sys.modules[pkg] = m
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref_root, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    parent = sys.modules.get(".".join(parts[:-1])) if len(parts) > 1 else None
    if parent is not None:
        setattr(parent, parts[-1], mod)
    return mod


def patch_reference(ref_root, voxel_mode=None):
    """Returns {reference name: replacement} for everything that was rebound."""
    ref_root = os.path.abspath(ref_root)
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    if voxel_mode is not None:
        os.environ["OPENESS_B200_VOXEL_MODE"] = voxel_mode

    from .datasets import data_util as b_du
    from .DSEC.dataset import representations as b_rep
    from .evaluation import metrics as b_met
    from .utils import loss_functions as b_lf

    done = {}

    def rebind(mod, names, src):
        for n in names:
            setattr(mod, n, getattr(src, n))
            done[f"{mod.__name__}.{n}"] = getattr(src, n)

    du = _load(ref_root, "datasets.data_util", "datasets/data_util.py")
    rebind(du, ["generate_input_representation", "generate_event_histogram", "normalize_voxel_grid",
                "generate_voxel_grid"], b_du)
    rep = _load(ref_root, "DSEC.dataset.representations", "DSEC/dataset/representations.py")
    rebind(rep, ["VoxelGrid"], b_rep)
    lf = _load(ref_root, "utils.loss_functions", "utils/loss_functions.py")
    rebind(lf, ["TaskLoss", "NCELoss", "DiceLoss", "symJSDivLoss"], b_lf)
    met = _load(ref_root, "evaluation.metrics", "evaluation/metrics.py")
    rebind(met, ["semseg_compute_confusion", "semseg_accum_confusion_to_iou", "semseg_accum_confusion_to_acc",
                 "MetricsSemseg"], b_met)
    try:   # needs cv2 / scipy, present in the reference's environment
        from .e2vid.utils import inference_utils as b_iu
        iu = _load(ref_root, "e2vid.utils.inference_utils", "e2vid/utils/inference_utils.py")
        rebind(iu, ["EventPreprocessor"], b_iu)
    except Exception as e:  # pragma: no cover - depends on the reference's optional imports
        done["e2vid.utils.inference_utils.EventPreprocessor"] = e
    # model mirrors (state_dict-compatible; tensor-core forward paths).  These reference modules import their
    # environment's optional packages (mmcv via models/__init__.py, cv2, ...): rebinding is best effort per module.
    def try_rebind(ref_name, rel, names, src_mod, stand_in=False):
        """stand_in: if the reference module cannot be imported (its third-party imports are missing), register the
        mirror module itself under the reference name -- `from models.maskclip_model import ...` then needs no mmcv."""
        try:
            src = __import__(src_mod, fromlist=["_"])
            try:
                mod = _load(ref_root, ref_name, rel)
            except ImportError:
                if not stand_in:
                    raise
                sys.modules.pop(ref_name, None)
                sys.modules[ref_name] = mod = src
                parent = sys.modules.get(ref_name.rpartition(".")[0])
                if parent is not None:
                    setattr(parent, ref_name.rpartition(".")[2], src)
            rebind(mod, names, src)
        except Exception as e:  # pragma: no cover - depends on the reference's optional imports
            for n in names:
                done[f"{ref_name}.{n}"] = e

    try_rebind("models.style_networks", "models/style_networks.py", ["SemSegE2VID"], "openess_b200.models.style_networks")
    try_rebind("models.image_model", "models/image_model.py", ["DilationFeatureExtractor"], "openess_b200.models.image_model")
    try_rebind("models.deeplabv3", "models/deeplabv3.py", ["deeplabv3_resnet50"], "openess_b200.models.deeplabv3")
    # e2vid.utils.loading_utils.load_model does `from e2vid.model.model import *` + eval(arch)(config): rebinding the class
    # in e2vid.model.model makes the reference's own checkpoint loader build the mirror (same state_dict keys).
    try_rebind("e2vid.model.model", "e2vid/model/model.py", ["E2VIDRecurrent"], "openess_b200.e2vid.model.model")
    try_rebind("e2vid.image_reconstructor", "e2vid/image_reconstructor.py", ["ImageReconstructor"],
               "openess_b200.e2vid.image_reconstructor")
    # rows a14 / a14' (modules the trainers construct or BASELINE config 2 names) and the DDD17 record access (8f row 1)
    try_rebind("models.maskclip_model", "models/maskclip_model.py",
               ["maskClipFeatureExtractor", "VisionTransformer", "TransformerEncoderLayer", "MaskClipHead", "PatchEmbed"],
               "openess_b200.models.maskclip_model", stand_in=True)
    try_rebind("models._resnet", "models/_resnet.py", ["ResNet", "resnet18", "resnet34", "resnet50"], "openess_b200.models._resnet")
    try_rebind("datasets.extract_data_tools.example_loader_ddd17", "datasets/extract_data_tools/example_loader_ddd17.py",
               ["load_files_in_directory", "load_events", "extract_events_from_memmap"],
               "openess_b200.datasets.extract_data_tools.example_loader_ddd17", stand_in=True)
    try_rebind("DSEC.utils.eventslicer", "DSEC/utils/eventslicer.py", ["EventSlicer"], "openess_b200.DSEC.utils.eventslicer",
               stand_in=True)
    # what `models/__init__.py` exports (`from models.image_model import *`, `from models.maskclip_model import *`), without
    # executing it: the trainers do `from models import Preprocessing, maskClipFeatureExtractor`
    models_pkg = sys.modules.get("models")
    if models_pkg is not None:
        for sub in ("models.image_model", "models.maskclip_model"):
            m = sys.modules.get(sub)
            if m is not None:
                for n in getattr(m, "__all__", [k for k in vars(m) if not k.startswith("_")]):
                    if not hasattr(models_pkg, n):
                        setattr(models_pkg, n, getattr(m, n))
    # zero-edit drop-in of the fused pretraining step and of device-side sample assembly (training/drop_in.py): best effort,
    # the trainer / dataset modules import their environment's optional packages (matplotlib, h5py, hdf5plugin, numba, ...)
    from .training import drop_in
    try:
        pt = _load(ref_root, "training.pretrain_trainer", "training/pretrain_trainer.py")
        done.update(drop_in.install(pretrain_trainer_module=pt))
    except Exception as e:  # pragma: no cover - depends on the reference's optional imports
        done["training.pretrain_trainer.OpenESSPretrainModel.task_train_step"] = e
    try:
        seq = _load(ref_root, "DSEC.dataset.sequence_ov", "DSEC/dataset/sequence_ov.py")
        done.update(drop_in.install(sequence_module=seq))
    except Exception as e:  # pragma: no cover
        done["DSEC.dataset.sequence_ov.Sequence.__getitem__"] = e
    _PATCHED.update(done)
    return done
