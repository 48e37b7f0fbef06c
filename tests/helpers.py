import torch


def make_paged(kv_lens, hkv, d, ps, layout="NHD", dtype=torch.float32, device="cpu", extra_pages=3):
    npages = [(l + ps - 1) // ps for l in kv_lens]
    total = sum(npages)
    indptr = torch.tensor([0] + torch.tensor(npages).cumsum(0).tolist(), dtype=torch.int32)
    indices = torch.randperm(total + extra_pages)[:total].int()
    last = torch.tensor([(l - 1) % ps + 1 if l > 0 else 0 for l in kv_lens], dtype=torch.int32)
    shape = (total + extra_pages, ps, hkv, d) if layout == "NHD" else (total + extra_pages, hkv, ps, d)
    kc = torch.randn(shape, device=device, dtype=dtype)
    vc = torch.randn(shape, device=device, dtype=dtype)
    return indptr, indices, last, kc, vc


def cos_sim(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return torch.nn.functional.cosine_similarity(a, b, dim=0).item()
