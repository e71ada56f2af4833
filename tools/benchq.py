import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], d["config"]["latents_per_launch"], "pipelined ms", round(d["ms_per_step"],4), "serial ms", round(d["roofline"]["step"]["serial_ms"],4), "embed", round(d["roofline"]["kernels"]["embed_kernel"]["ms"],4), "extract", round(d["roofline"]["kernels"]["extract_kernel"]["ms"],4), d["config"]["decode_exact"])
