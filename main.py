"""Entry point with the reference's CLI surface (main.py:29-36):
    python main.py [--method M] [--backbone B] [--dem] [--slope] [--batch_size N] [--inputs a b] [--seed S]
Config merge order follows main.py:48-69: config.json <- method/<m>/<m>.json <- data_config <- train_config."""
import argparse
import os
import pprint
import random
from datetime import datetime
from pathlib import Path

import numpy as np
import torch

from kurosiwo_b200 import json5lite, synthetic
from kurosiwo_b200.change_detection_trainer import eval_change_detection, train_change_detection
from kurosiwo_b200.model_utilities import initialize_cd_model, initialize_segmentation_model
from kurosiwo_b200.segmentation_trainer import eval_semantic_segmentation, train_semantic_segmentation
from kurosiwo_b200.utilities import update_config

parser = argparse.ArgumentParser()
parser.add_argument("--method", default=None)
parser.add_argument("--backbone", default=None)
parser.add_argument("--dem", action="store_true", default=False)
parser.add_argument("--slope", action="store_true", default=False)
parser.add_argument("--batch_size", default=None)
parser.add_argument("--inputs", nargs="+", default=None)
parser.add_argument("--seed", type=int, default=999)


def load_configs(args, root=Path(__file__).resolve().parent):
    configs = json5lite.load(open(root / "configs/config.json"))
    if args.method is not None:
        configs["method"] = args.method
    m = configs["method"].lower()
    if m == "finetune":
        configs["task"] = "segmentation"                   # the reference selects the trainer by task (main.py:140-200)
        if args.inputs is None:
            configs["inputs_override"] = ["pre_event_1", "pre_event_2", "post_event"]
    model_configs = json5lite.load(open(root / f'configs/method/{m}/{m.replace("-", "_")}.json'))
    if args.backbone is not None:
        model_configs["backbone"] = args.backbone          # stored and ignored by CD models, as in the reference
    configs.update(model_configs)
    configs.update(json5lite.load(open(root / "configs/train/data_config.json")))
    configs.update(json5lite.load(open(root / "configs/train/train_config.json")))
    if args.inputs is not None:
        configs["inputs"] = args.inputs
    elif "inputs_override" in configs:
        configs["inputs"] = configs.pop("inputs_override")
    if args.dem:
        configs["dem"] = True
    configs = update_config(configs, args)
    if args.batch_size is not None:
        configs["batch_size"] = int(args.batch_size)
    run_ts = datetime.now().strftime("%Y%m%d%H%M%S")
    configs["checkpoint_path"] = f'checkpoints/{m}/{configs["track"]}_{run_ts}'
    return configs, model_configs


def prepare_loaders(configs):
    if not configs.get("synthetic", False):
        raise NotImplementedError("the GeoTIFF dataset pipeline (dataset/Dataset.py) is outside the B200 hot path; set synthetic=true")
    nb = configs.get("synthetic_batches", 4)
    mk = lambda seed, n: synthetic.SyntheticLoader(configs["batch_size"], n, seed=seed)
    return mk(configs["seed"], nb), mk(configs["seed"] + 100, max(1, nb // 2)), mk(configs["seed"] + 200, max(1, nb // 2))


def init_data_parallel(configs):
    """One process per GPU under `torchrun` (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the environment): NCCL over NVLink only.
    The reference is single-process; a plain `python main.py` run stays that way (returns None)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    import torch.distributed as dist
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")            # gradient all-reduce over NVLink / NVSwitch peer access only
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if configs.get("gpu") is not None:
        configs["gpu"] = local
        configs["device"] = f"cuda:{local}"
        torch.cuda.set_device(local)
    dist.init_process_group("nccl" if configs.get("gpu") is not None else "gloo")
    configs["seed"] = int(configs.get("seed", 999)) + dist.get_rank()   # every rank draws its own batches
    return dist.group.WORLD


def load_best_checkpoint(model, configs, task):
    """The reference evaluates the test set on `best_segmentation.pt`, not on the last-epoch weights (main.py:149-152: the pickled
    module for segmentation; :172-185: `checkpoint["model_state_dict"]` for change detection).  Falls back to the current weights
    when no best checkpoint was written (e.g. `test: true` on a fresh directory, or no validation loader)."""
    path = Path(configs["checkpoint_path"]) / "best_segmentation.pt"
    if not path.exists():
        print(f"No best checkpoint at {path}: evaluating the current weights")
        return model
    print("Loading model from: ", str(path))
    from kurosiwo_b200.checkpoint_compat import load_reference_checkpoint   # also accepts files written by the reference itself
    return load_reference_checkpoint(path, model, configs["device"], getattr(model, "precision", "bf16"), configs)


if __name__ == "__main__":
    args = parser.parse_args()
    np.random.seed(args.seed); random.seed(args.seed); torch.manual_seed(args.seed)
    configs, model_configs = load_configs(args)
    configs["seed"] = args.seed
    pg = init_data_parallel(configs)
    rank0 = pg is None or torch.distributed.get_rank() == 0
    if pg is not None:      # every rank must agree on the checkpoint directory (the name carries a timestamp)
        box = [configs["checkpoint_path"]]
        torch.distributed.broadcast_object_list(box, src=0)
        configs["checkpoint_path"] = box[0]
    if rank0:
        Path(configs["checkpoint_path"]).mkdir(parents=True, exist_ok=True)
        pprint.pprint(configs)
    train_loader, val_loader, test_loader = prepare_loaders(configs)
    if configs["task"] == "segmentation":
        model = initialize_segmentation_model(configs, model_configs).to(configs["device"])
        if not configs["test"]:
            train_semantic_segmentation(model, train_loader, val_loader, test_loader, configs=configs, model_configs=model_configs, process_group=pg)
        if pg is not None:
            torch.distributed.barrier()
        model = load_best_checkpoint(model, configs, "segmentation")
        test_acc, test_score, miou = eval_semantic_segmentation(model, test_loader, configs=configs, settype="Test", model_configs=model_configs)
        if rank0:
            print("Test Mean IOU: ", miou)
    elif configs["task"] == "cd":
        model = initialize_cd_model(configs, model_configs)
        if not configs["test"]:
            train_change_detection(model, train_loader, val_loader, test_loader, configs=configs, model_configs=model_configs, process_group=pg)
        if pg is not None:
            torch.distributed.barrier()
        model = load_best_checkpoint(model, configs, "cd")
        test_acc, test_score, miou = eval_change_detection(model, test_loader, settype="Test", configs=configs, model_configs=model_configs)
        if rank0:
            print("Test Mean IOU: ", miou.item())
    else:
        raise NotImplementedError("tasks 'cd' and 'segmentation' are on the B200 hot path (SURVEY.md §8); 'mae' pre-training is not")
    if pg is not None:
        torch.distributed.destroy_process_group()
