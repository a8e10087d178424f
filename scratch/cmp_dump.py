import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
for k in a.files:
    x, y = a[k], b[k]
    den = np.abs(x).max()
    print(k, "max abs", float(np.abs(x - y).max()), "rel", float(np.abs(x - y).max() / den), "nan", int(np.isnan(y).sum()))
r = np.abs(a["Ksum"] - b["Ksum"]) / np.abs(a["Ksum"])
print("worst trajectories", np.argsort(-r)[:5], np.sort(r)[-5:])
