"""Pinned host<->device copy bandwidth of the box (one direction alone, both at once): the ceiling of bench.py's e2e leg."""
import json
import torch

def main():
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        for it in range(2):
            torch.cuda.synchronize()
            e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            s1.wait_event(e0); s2.wait_event(e0)
            for _ in range(8):
                if up:
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_in, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_b, non_blocking=True)
            e1.record(s1); e2.record(s2)
            torch.cuda.synchronize()
            ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
        out[name] = {"GBs_per_direction": 8 * n / ms / 1e6}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
