"""OCaml `result` look-alikes so the parity tests read like the reference's:
    assert De.Inf.Ns.inflate(src, dst) == Ok((6, 1))
    assert De.Inf.Ns.inflate(src, dst) == Error("Invalid_complement_of_length")
"""


class Ok:
    __slots__ = ("value",)

    def __init__(self, value=None):
        self.value = value

    def __eq__(self, other):
        return isinstance(other, Ok) and self.value == other.value

    def __repr__(self):
        return "Ok(%r)" % (self.value,)

    def is_ok(self):
        return True


class Error:
    __slots__ = ("error",)

    def __init__(self, error):
        self.error = error

    def __eq__(self, other):
        return isinstance(other, Error) and self.error == other.error

    def __repr__(self):
        return "Error(%r)" % (self.error,)

    def is_ok(self):
        return False
