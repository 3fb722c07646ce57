import os, sys
print("hello from rank", os.environ.get("RANK"), "local", os.environ.get("LOCAL_RANK"), flush=True)
import torch
print("cuda devices", torch.cuda.device_count(), flush=True)
