"""Per-op CUDA-event timing of one eager training step of a bench workload:  python scripts/prof_ops.py <workload> [batch] [opt=value,...]"""
import collections, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from kurosiwo_b200 import synthetic
wl_name = sys.argv[1]; wl = bench.WORKLOADS[wl_name]; bs = int(sys.argv[2]) if len(sys.argv) > 2 else wl["batch"]
dev = "cuda:0"
torch.manual_seed(0)
configs = {"device": dev, "inputs": ["pre_event_1", "post_event"], "dem": False, "scale_input": "normalize", "num_classes": 3, "num_channels": 2,
           "loss_function": "ce+dice", "class_weights": [1.0, 1.0, 1.0], "method": wl["method"], "epochs": 1, "precision": "bf16", "resume_checkpoint": False}
mc = {"method": wl["method"], "optimizer": "adam", "learning_rate": wl["lr"], "lr_schedule": None, "base_channel": 32, "embed_dim": 256, "decoder_softmax": True}
b = synthetic.make_batch(999, bs)
if wl["task"] == "cd":
    from kurosiwo_b200.model_utilities import initialize_cd_model
    from kurosiwo_b200.change_detection_trainer import FusedStepper
    if wl["method"] == "changeformer":
        mc.update({"optimizer": "sgd", "momentum": 0.99, "weight_decay": 1e-5})
    model = initialize_cd_model(configs, mc).train()
    stepper = FusedStepper(model, configs, mc)
    inputs = (b[6].to(dev), b[2].to(dev), b[3].to(dev))
else:
    from kurosiwo_b200.model_utilities import initialize_segmentation_model
    from kurosiwo_b200.segmentation_trainer import FusedSegStepper
    configs.update({"inputs": ["pre_event_1", "pre_event_2", "post_event"], "num_channels": 6, "mlp": False, "decoder": False, "task": "segmentation",
                    "finetuning_patch_size": 16, "linear_eval": False, "encoder": None})
    mc["encoder_config"] = {"image_size": 224, "patch_size": 16, "dim": 768, "depth": 12, "heads": 12, "mlp_dim": 3072}
    if wl_name == "floodvit-upernet":
        configs["head"] = "upernet"
    model = initialize_segmentation_model(configs, mc).to(dev).train()
    stepper = FusedSegStepper(model, configs, mc)
    inputs = (torch.cat((b[2], b[6], b[9]), 1).to(dev), b[3].to(dev))
eng = stepper._engine(inputs[0]); ops = eng.ops
for kv in (sys.argv[3].split(',') if len(sys.argv) > 3 else []):
    ops.set_option(kv.split('=')[0], int(kv.split('=')[1]))
for _ in range(2): eng.train_step(*inputs)
torch.cuda.synchronize()
times = collections.OrderedDict(); orig = {}
def wrap(name):
    f = getattr(ops, name)
    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **k); e1.record()
        key = name
        if name in ("conv2d", "conv2d_wgrad"): key = f"{name}[k{a[3]}]" + ("[K>=1024]" if a[3] == 1 and sum(v.C for v in a[4]) >= 1024 else "")
        if name.startswith("conv2d_strided"): key = f"{name}[k{a[5]}s{a[6]}]"
        times.setdefault(key, []).append((e0, e1)); return r
    orig[name] = f; setattr(ops, name, g)
for n in [m for m in dir(ops) if not m.startswith("_") and callable(getattr(ops, m)) and m not in ("set_option", "make_permute_table", "ce_dice_workspace")]:
    wrap(n)
eng.train_step(*inputs); torch.cuda.synchronize()
for n, f in orig.items(): setattr(ops, n, f)
rows = sorted(((sum(a.elapsed_time(b_) for a, b_ in ev), n, len(ev)) for n, ev in times.items()), reverse=True)
tot = sum(r[0] for r in rows)
for ms, n, c in rows: print(f"{n:32s} n={c:4d} {ms:9.3f} ms {100 * ms / tot:5.1f}%")
print("sum", tot)
