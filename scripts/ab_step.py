"""A/B of library options on the CUDA-graph training step of a bench workload, in ONE process on ONE GPU (box-to-box variance is ~3 %):
   python scripts/ab_step.py snunet stem_simt=1 ecam_simt=1 "loss_variant=1,stem_simt=1"
Prints ms/step for the default options and for every listed option set (comma-separated name=value pairs)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from kurosiwo_b200 import synthetic
wl_name = sys.argv[1]; wl = bench.WORKLOADS[wl_name]; bs = wl["batch"]
dev = "cuda:0"
torch.manual_seed(0)
configs = {"device": dev, "inputs": ["pre_event_1", "post_event"], "dem": False, "scale_input": "normalize", "num_classes": 3, "num_channels": 2,
           "loss_function": "ce+dice", "class_weights": [1.0, 1.0, 1.0], "method": wl["method"], "epochs": 1, "precision": "bf16", "resume_checkpoint": False}
mc = {"method": wl["method"], "optimizer": "adam", "learning_rate": wl["lr"], "lr_schedule": None, "base_channel": 32, "embed_dim": 256, "decoder_softmax": True}
b = synthetic.make_batch(999, bs)
assert wl["task"] == "cd"
from kurosiwo_b200.model_utilities import initialize_cd_model
from kurosiwo_b200.change_detection_trainer import FusedStepper
if wl["method"] == "changeformer":
    mc.update({"optimizer": "sgd", "momentum": 0.99, "weight_decay": 1e-5})
model = initialize_cd_model(configs, mc).train()
stepper = FusedStepper(model, configs, mc)
inputs = (b[6].to(dev), b[2].to(dev), b[3].to(dev))
eng = stepper._engine(inputs[0]); ops = eng.ops


def measure(opts, reps=12):
    for k, v in opts.items():
        ops.set_option(k, v)
    try:
        eng.graph = None
        eng.capture(*inputs, warmup=2)
        for _ in range(3):
            eng.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    finally:
        for k in opts:
            ops.set_option(k, 0)


sets = [{}] + [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in a.split(",")) for a in sys.argv[2:]] + [{}]
for o in sets:
    print(f"{wl_name} bs={bs} {o or 'default'}: {measure(o):.3f} ms/step", flush=True)
