class KeyBindings:
    def add(self, *a, **k):
        return lambda f: f
