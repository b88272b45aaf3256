"""single-rank MPI stand-in (see package docstring)"""
import numpy as np

IN_PLACE = object()
MODE_RDONLY = MODE_WRONLY = MODE_CREATE = 0
DOUBLE = INT64_T = object()


class _Request:
    @staticmethod
    def Waitall(requests):
        return None

    def Wait(self):
        return None


Request = _Request


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Allreduce(self, sendbuf, recvbuf, op=None):
        if sendbuf is not IN_PLACE:
            np.asarray(recvbuf)[...] = np.asarray(sendbuf)

    def Iallreduce(self, sendbuf, recvbuf, op=None):
        self.Allreduce(sendbuf, recvbuf)
        return _Request()

    def Allgather(self, sendbuf, recvbuf):
        np.asarray(recvbuf).reshape(-1)[...] = np.asarray(sendbuf).reshape(-1)

    def Bcast(self, buf, root=0):
        return None

    def Ibcast(self, buf, root=0):
        return _Request()

    def bcast(self, obj, root=0):
        return obj

    def allreduce(self, obj, op=None):
        return obj

    def Barrier(self):
        return None

    def barrier(self):
        return None


COMM_WORLD = _Comm()
SUM = object()
