import sys, json, os
sys.path.insert(0, "/root/repo")
import torch, bench
from ml_accelerated_simulation_b200 import models
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
print(bench._config1_gpu(model, torch.device("cuda:0")))
