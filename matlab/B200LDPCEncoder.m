%B200LDPCENCODER Drop-in for comm.LDPCEncoder at NRLDPCEncoder.m:49 backed by libnrldpc_b200.
%   UNVERIFIED (no MATLAB in the build image) -- see INTEGRATION.md.
%
%   obj.hLDPCEncoder = B200LDPCEncoder('ParityCheckMatrix',obj.H);
%   cw = step(obj.hLDPCEncoder, c);                          % NRLDPCEncoder.m:158 unchanged
classdef B200LDPCEncoder < matlab.System
    properties (Nontunable)
        ParityCheckMatrix
    end
    properties (Access = private)
        handle = uint64(0);
    end
    methods
        function obj = B200LDPCEncoder(varargin)
            setProperties(obj, nargin, varargin{:});
        end
        function delete(obj)
            releaseImpl(obj);
        end
    end
    methods (Access = protected)
        function setupImpl(obj)
            [m, n] = size(obj.ParityCheckMatrix);
            if mod(n,68) == 0 && m == 46*n/68
                BG = 1; Z = n/68;
            elseif mod(n,52) == 0 && m == 42*n/52
                BG = 2; Z = n/52;
            else
                error('ldpc_3gpp_matlab:UnsupportedParameters','H is not a 3GPP NR parity check matrix.');
            end
            obj.handle = nrldpc_mex('create', BG, Z, 1, 0, 0.75);
        end
        function cw = stepImpl(obj, c)
            cw = nrldpc_mex('encode', obj.handle, c);
        end
        function releaseImpl(obj)
            if obj.handle ~= 0
                nrldpc_mex('destroy', obj.handle);
                obj.handle = uint64(0);
            end
        end
    end
end
