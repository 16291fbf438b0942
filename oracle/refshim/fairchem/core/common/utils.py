from functools import wraps


def conditional_grad(dec):
    """Apply `dec` iff self.regress_forces and not self.direct_forces (fairchem semantics)."""
    def decorator(func):
        @wraps(func)
        def cls_method(self, *args, **kwargs):
            f = func
            if getattr(self, "regress_forces", False) and not getattr(self, "direct_forces", 0):
                f = dec(func)
            return f(self, *args, **kwargs)
        return cls_method
    return decorator
