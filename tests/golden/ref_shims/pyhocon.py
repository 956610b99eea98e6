class ConfigFactory:
    pass
