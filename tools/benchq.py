import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"])); [print("  ",k, round(v["avg_launch_ms"],4), round(v["frac"],3)) for k,v in d["roofline_kernels"].items()]
