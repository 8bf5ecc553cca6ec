// filled in with the QP oracle
