def stub(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(f"dropin_tree: {name} is a stub -- excel_b200.install() did not patch it")
    fn.__name__ = name.rsplit(".", 1)[-1]
    return fn
