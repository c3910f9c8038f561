import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaired_image_generation_b200 as cgb
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mods = (cgb.Generator(seed=1), cgb.Generator(seed=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
tr = cgb.CycleGANTrainer(*mods)
a = (torch.rand(batch, 3, 256, 256) * 2 - 1).cuda(); b = (torch.rand(batch, 3, 256, 256) * 2 - 1).cuda()
eng = tr._ensure_engine(a)
with torch.cuda.stream(tr.stream):
    eng.set_inputs(a, b)
    print(eng.timeline())
