"""Drop-in package for `simple_knn` (camenduru/simple-knn @ 60f461f4); see simple_knn._C."""
