"""Entry point with the reference's CLI surface (main.py:29-36):
    python main.py [--method M] [--backbone B] [--dem] [--slope] [--batch_size N] [--inputs a b] [--seed S]
Config merge order follows main.py:48-69: config.json <- method/<m>/<m>.json <- data_config <- train_config."""
import argparse
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


if __name__ == "__main__":
    args = parser.parse_args()
    np.random.seed(args.seed); random.seed(args.seed); torch.manual_seed(args.seed)
    configs, model_configs = load_configs(args)
    configs["seed"] = args.seed
    Path(configs["checkpoint_path"]).mkdir(parents=True, exist_ok=True)
    pprint.pprint(configs)
    train_loader, val_loader, test_loader = prepare_loaders(configs)
    if configs["task"] == "segmentation":
        model = initialize_segmentation_model(configs, model_configs).to(configs["device"])
        if not configs["test"]:
            train_semantic_segmentation(model, train_loader, val_loader, test_loader, configs=configs, model_configs=model_configs)
        test_acc, test_score, miou = eval_semantic_segmentation(model, test_loader, configs=configs, settype="Test", model_configs=model_configs)
        print("Test Mean IOU: ", miou)
        raise SystemExit(0)
    if configs["task"] != "cd":
        raise NotImplementedError("tasks 'cd' and 'segmentation' are on the B200 hot path (SURVEY.md §8); 'mae' pre-training is not")
    model = initialize_cd_model(configs, model_configs)
    if not configs["test"]:
        train_change_detection(model, train_loader, val_loader, test_loader, configs=configs, model_configs=model_configs)
    test_acc, test_score, miou = eval_change_detection(model, test_loader, settype="Test", configs=configs, model_configs=model_configs)
    print("Test Mean IOU: ", miou)
