"""Empty stand-in (rendering is never called by the oracle)."""


class WebsocketServer:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError
