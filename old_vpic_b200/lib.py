"""placeholder"""
def load():
    raise RuntimeError("not built")
