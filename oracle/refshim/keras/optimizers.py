Adam = object
