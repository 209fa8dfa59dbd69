rcParams = {"text.usetex": False}
