import os, sys
print("hello from rank", os.environ.get("RANK"), "local", os.environ.get("LOCAL_RANK"), flush=True)
import torch
print("cuda", torch.cuda.is_available(), torch.cuda.device_count(), flush=True)
