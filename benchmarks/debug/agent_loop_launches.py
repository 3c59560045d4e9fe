"""One replay of the captured Test_Agent loop (environment.capture_rollout with the reference's agent as policy) for
`ncu --metrics gpu__time_duration.sum`: which kernels an iteration at batch B consists of.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python benchmarks/debug/agent_loop_launches.py 1"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import cmr_agent_b200
from cmr_agent_b200 import agent_tower, environment as env, synth
from oracle import reference_loader as rl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
rl.put_on_path()
cmr_agent_b200.install()
from config import KittiConfiguration
from models import CMRAgent
config = KittiConfiguration()
torch.manual_seed(2023)
agent = agent_tower.accelerate_agent(CMRAgent(config).to(dev).eval())
cpu = synth.make_batch(B, first_episode=3, seed=2023)
data = dict(cpu)
for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"):
    data[k] = cpu[k].to(dev)
torch.distributions.Distribution.set_default_validate_args(False)
with torch.no_grad():
    roll = env.capture_rollout(data, config, with_reward=False, reusable=True,
                               policy=lambda s2, s3: agent.action_from_logits(*agent(s2, s3)[:2], deterministic=True))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    roll.replay()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
