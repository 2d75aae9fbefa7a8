%B200LDPCDECODER Drop-in for comm.LDPCDecoder at NRLDPCDecoder.m:120 backed by libnrldpc_b200.
%   UNVERIFIED (no MATLAB in the build image) -- see INTEGRATION.md.
%
%   obj.hLDPCDecoder = B200LDPCDecoder('ParityCheckMatrix',obj.H, ...
%       'MaximumIterationCount',obj.iterations, ...
%       'IterationTerminationCondition','Parity check satisfied');
%   c_hat = double(step(obj.hLDPCDecoder, cw_tilde));        % NRLDPCDecoder.m:265 unchanged
%
%   (BG, Z) are inferred from size(H) = 46Z x 68Z or 42Z x 52Z and H is verified against the
%   library's own lifted table, so a matrix that is not a TS 38.212 matrix is rejected.
classdef B200LDPCDecoder < matlab.System
    properties (Nontunable)
        ParityCheckMatrix
        MaximumIterationCount = 50;
        IterationTerminationCondition = 'Maximum iteration count';
        Alpha = 0.75;          % min-sum normalisation (engine-specific)
        ActiveRows = 0;        % 0 = all base rows (engine-specific)
        Algorithm = 'Layered normalized min-sum';  % or 'Sum-product': comm.LDPCDecoder's own flooding sum-product in
                                                   % float64 (matches the CPU restatement of the documented algorithm;
                                                   % against the toolbox itself: run make_golden_vectors.m)
    end
    properties (Access = private)
        handle = uint64(0);
    end
    methods
        function obj = B200LDPCDecoder(varargin)
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
            if ~isequal(obj.ParityCheckMatrix, get_pcm(get_3gpp_base_graph(BG, get_3gpp_set_index(Z)), Z))
                error('ldpc_3gpp_matlab:UnsupportedParameters','H does not match TS 38.212 for BG%d, Z=%d.', BG, Z);
            end
            early = strcmp(obj.IterationTerminationCondition, 'Parity check satisfied');
            alg = double(strcmp(obj.Algorithm, 'Sum-product'));
            obj.handle = nrldpc_mex('create', BG, Z, obj.MaximumIterationCount, double(early), obj.Alpha, 0, alg);
        end
        function c_hat = stepImpl(obj, cw_tilde)
            c_hat = nrldpc_mex('decode', obj.handle, double(cw_tilde), obj.ActiveRows);
        end
        function releaseImpl(obj)
            if obj.handle ~= 0
                nrldpc_mex('destroy', obj.handle);
                obj.handle = uint64(0);
            end
        end
    end
end
