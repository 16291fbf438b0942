class BackboneInterface:
    pass
