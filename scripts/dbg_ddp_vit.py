"""2-GPU debug of the bucketed all-reduce path (torchrun).  Prints stage by stage."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
def say(*a):
    print(f"[r{rank} {time.time()%1000:.1f}]", *a, flush=True)
from kurosiwo_b200.vision_transformer import FinetunerSegmentation, ViT
torch.manual_seed(0)
enc = ViT(image_size=224, patch_size=16, num_classes=3, dim=256, depth=4, heads=4, mlp_dim=512, channels=6, precision="bf16")
model = FinetunerSegmentation(encoder=enc, configs={"mlp": False, "decoder": False, "num_classes": 3, "finetuning_patch_size": 16}).to(dev).train()
img = torch.randn(8, 6, 224, 224, device=dev); mask = torch.randint(0, 4, (8, 224, 224), device=dev)
eng = model.engine(img)
eng.init_training(lr=1e-4, process_group=dist.group.WORLD, bucket_mb=float(os.environ.get("BUCKET_MB", "2")))
say("init done")
for i in range(2):
    eng.train_step(img, mask); torch.cuda.synchronize(); say("eager step", i, eng.comm_stats)
step = eng.capture(img, mask, warmup=0)
torch.cuda.synchronize(); say("captured", [(k, (len(x) if k == "comm" else "g")) for k, x in eng.graph])
for i in range(3):
    step(); torch.cuda.synchronize(); say("replay", i, float(eng.loss3[0]))
ref = eng.params.flat.clone()
dist.all_reduce(ref); 
say("replica check", float((ref / 2 - eng.params.flat).abs().max()))
dist.destroy_process_group()
