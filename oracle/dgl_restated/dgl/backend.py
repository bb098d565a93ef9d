def asnumpy(t):
    return t.detach().cpu().numpy()
