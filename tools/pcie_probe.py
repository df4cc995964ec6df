import torch, time
dev=torch.device("cuda",0)
def bw(nbytes_h2d, nbytes_d2h, chunks):
    h_in = torch.empty(nbytes_h2d, dtype=torch.uint8, pin_memory=True) if nbytes_h2d else None
    d_in = torch.empty(nbytes_h2d, dtype=torch.uint8, device=dev) if nbytes_h2d else None
    h_out = torch.empty(nbytes_d2h, dtype=torch.uint8, pin_memory=True) if nbytes_d2h else None
    d_out = torch.empty(nbytes_d2h, dtype=torch.uint8, device=dev) if nbytes_d2h else None
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run():
        if h_in is not None:
            with torch.cuda.stream(s1):
                for c in range(chunks):
                    n = nbytes_h2d // chunks
                    d_in[c*n:(c+1)*n].copy_(h_in[c*n:(c+1)*n], non_blocking=True)
        if h_out is not None:
            with torch.cuda.stream(s2):
                for c in range(chunks):
                    n = nbytes_d2h // chunks
                    h_out[c*n:(c+1)*n].copy_(d_out[c*n:(c+1)*n], non_blocking=True)
    for _ in range(3): run()
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(10): run()
    torch.cuda.synchronize()
    dt=(time.perf_counter()-t0)/10
    return dt*1e3, (nbytes_h2d/dt/1e9 if nbytes_h2d else 0), (nbytes_d2h/dt/1e9 if nbytes_d2h else 0)
A, B = 100663296, 67108864
for ch in (1, 8, 32):
    print("chunks", ch, "h2d only ms/GBs", bw(A,0,ch), "d2h only", bw(0,B,ch), "both", bw(A,B,ch))
