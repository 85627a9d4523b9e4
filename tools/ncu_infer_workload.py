import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import fpga_cnn_b200 as fc, inputs
wt = np.fromfile("tests/golden/weights.bin", dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt); acc.set_shifts(2, 4, 6)
fw, fb = inputs.make_fc(); acc.load_classifier(fw, fb)
x = torch.randint(0, 256, (16384, 128, 128), dtype=torch.uint8, device="cuda")
for i in range(4):
    acc.infer_batch(x)
torch.cuda.synchronize()
