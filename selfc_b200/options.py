"""Options loader with the reference's behaviour (codes/options/options.py:9-102): YAML -> ordered dict with the
derived `is_train`, per-dataset `phase` / `data_type`, `path` entries, then `dict_to_nonedict` so missing keys
read as None.  The unmodified reference YAMLs (e.g. options/test/rescaling/test_SelfC_large_vid4.yml) drive it.

One deliberate difference: the reference exports CUDA_VISIBLE_DEVICES from `gpu_ids`
(options.py:13-16), which on an 8-GPU box pins every rank to the YAML's GPU.  Here that export only happens when
`honour_gpu_ids=True` (or SELFC_B200_HONOUR_GPU_IDS=1); by default the launcher's device assignment wins.
"""
from __future__ import annotations

import os
import os.path as osp
from collections import OrderedDict

import yaml


def _ordered_loader():
    class Loader(yaml.SafeLoader):
        pass

    def construct_mapping(loader, node):
        loader.flatten_mapping(node)
        return OrderedDict(loader.construct_pairs(node))

    Loader.add_constructor(yaml.resolver.BaseResolver.DEFAULT_MAPPING_TAG, construct_mapping)
    return Loader


def parse(opt_path, is_train=True, honour_gpu_ids=None):
    with open(opt_path, mode="r") as f:
        opt = yaml.load(f, Loader=_ordered_loader())
    if honour_gpu_ids is None:
        honour_gpu_ids = os.environ.get("SELFC_B200_HONOUR_GPU_IDS", "0") == "1"
    if opt.get("gpu_ids") and honour_gpu_ids:
        gpu_list = ",".join(str(x) for x in opt["gpu_ids"])
        os.environ["CUDA_VISIBLE_DEVICES"] = gpu_list
        print("export CUDA_VISIBLE_DEVICES=" + gpu_list)

    opt["is_train"] = is_train
    scale = opt.get("scale")
    for phase, dataset in (opt.get("datasets") or {}).items():
        dataset["phase"] = phase.split("_")[0]
        if opt.get("distortion") == "sr":
            dataset["scale"] = scale
        is_lmdb = False
        for key in ("dataroot_GT", "dataroot_LQ"):
            if dataset.get(key) is not None:
                dataset[key] = osp.expanduser(dataset[key])
                is_lmdb = is_lmdb or dataset[key].endswith("lmdb")
        dataset["data_type"] = "lmdb" if is_lmdb else "img"
        if dataset.get("mode", "").endswith("mc"):
            dataset["data_type"] = "mc"
            dataset["mode"] = dataset["mode"].replace("_mc", "")

    opt.setdefault("path", OrderedDict())
    for key, path in opt["path"].items():
        if path and key != "strict_load" and isinstance(path, str):
            opt["path"][key] = osp.expanduser(path)
    root = osp.abspath(osp.join(osp.dirname(osp.abspath(__file__)), osp.pardir))
    opt["path"]["root"] = root
    if is_train:
        exp = osp.join(root, "experiments", opt["name"])
        opt["path"]["experiments_root"] = exp
        opt["path"]["models"] = osp.join(exp, "models")
        opt["path"]["training_state"] = osp.join(exp, "training_state")
        opt["path"]["log"] = exp
        opt["path"]["val_images"] = osp.join(exp, "val_images")
        if "debug" in opt["name"]:
            opt["train"]["val_freq"] = 8
            opt["logger"]["print_freq"] = 1
            opt["logger"]["save_checkpoint_freq"] = 8
    else:
        res = osp.join(root, "results", opt["name"])
        opt["path"]["results_root"] = res
        opt["path"]["log"] = res
    if opt.get("distortion") == "sr":
        opt["network_G"]["scale"] = scale
    return opt


class NoneDict(dict):
    def __missing__(self, key):
        return None


def dict_to_nonedict(opt):
    if isinstance(opt, dict):
        return NoneDict(**{k: dict_to_nonedict(v) for k, v in opt.items()})
    if isinstance(opt, list):
        return [dict_to_nonedict(v) for v in opt]
    return opt


def dict2str(opt, indent_l=1):
    msg = ""
    for k, v in opt.items():
        if isinstance(v, dict):
            msg += " " * (indent_l * 2) + k + ":[\n" + dict2str(v, indent_l + 1) + " " * (indent_l * 2) + "]\n"
        else:
            msg += " " * (indent_l * 2) + k + ": " + str(v) + "\n"
    return msg


def check_resume(opt, resume_iter):
    """options.py:105-120 of the reference (called at train.py:122): when resuming, the network weights are those saved next to
    the training state -- path.pretrain_model_G is pointed at <models>/<iter>_G.pth (a configured pretrain path is ignored,
    with a warning).  Unlike the reference this fails at once when that file is missing instead of training on from freshly
    initialised weights with stale Adam moments."""
    import logging
    import os.path as osp
    logger = logging.getLogger("base")
    if opt["path"].get("resume_state"):
        if opt["path"].get("pretrain_model_G") is not None:
            logger.warning("pretrain_model path will be ignored when resuming training.")
        path = osp.join(opt["path"]["models"], "{}_G.pth".format(resume_iter))
        if not osp.isfile(path):
            raise FileNotFoundError(f"resume_state is set but the matching weights {path} do not exist")
        opt["path"]["pretrain_model_G"] = path
        logger.info("Set [pretrain_model_G] to " + path)
